#!/bin/bash
# round-2 GPU call K: k-block rotation and main-chain priority A/B, probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2k_pytest.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2k_bench_$name.json 2> gpurun_out/r2k_bench_$name.err; }
b default X=1
b no_rotate FIRA_GEMM_ROTATE=0
b no_priority FIRA_MAIN_PRIORITY=0
b side4 FIRA_SIDE_STREAMS=4
b side16 FIRA_SIDE_STREAMS=16
timeout 300 python tools/gemm_probe.py > gpurun_out/r2k_gemm_probe.jsonl 2> gpurun_out/r2k_gemm_probe.err
timeout 600 python bench.py --steps 10 --warmup 5 --timeline gpurun_out/r2k_timeline.json > gpurun_out/r2k_timeline.log 2>&1
tail -n 5 gpurun_out/r2k_pytest.log
python - <<'PY'
import json
for n in ['default','no_rotate','no_priority','side4','side16']:
    try:
        for l in open(f'gpurun_out/r2k_bench_{n}.json'):
            if l.startswith('{'):
                d=json.loads(l); print(n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))
    except Exception as e: print(n,'ERR',e)
for l in open('gpurun_out/r2k_gemm_probe.jsonl'):
    d=json.loads(l); print(d['shape'], d['pdl'], d['chain_us_per_launch_median'], d['cta0_phase_ns'])
PY
