#!/bin/bash
# round-2 GPU call J: TMA-store epilogue of the tcgen05 GEMM, bias gradients folded into the wgrad GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2j_pytest_tc.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2j_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2j_bench_$name.json 2> gpurun_out/r2j_bench_$name.err; }
b default X=1
b no_tma_store FIRA_GEMM_TMA_STORE=0
b no_dbias FIRA_DBIAS_FUSED=0
timeout 300 python tools/gemm_probe.py > gpurun_out/r2j_gemm_probe.jsonl 2> gpurun_out/r2j_gemm_probe.err
for f in gpurun_out/r2j_pytest_*.log; do echo "== $f"; tail -n 12 $f; done
for n in default no_tma_store no_dbias; do head -c 160 gpurun_out/r2j_bench_$n.json | cut -c40-160; echo; tail -n 3 gpurun_out/r2j_bench_$n.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r2j_gemm_probe.jsonl'):
    d=json.loads(l); print(d['shape'], d['pdl'], d['chain_us_per_launch_median'], d['cta0_phase_ns'])
PY
