#!/bin/bash
# round-2 GPU call E (first call of the re-created container): the whole -m gpu suite the way the driver runs it,
# bench lines of the candidate configurations, single-kernel A/B timings, launch list of one bench step, ncu --set full
# captures of the tensor-core kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2e_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 $EXTRA > gpurun_out/r2e_bench_$name.json 2> gpurun_out/r2e_bench_$name.err; }
EXTRA="" b default X=1
EXTRA="--skip-cpu-baseline --skip-parity-mode" b fused FIRA_GCN_FUSED=1
EXTRA="--skip-cpu-baseline --skip-parity-mode" b overlap FIRA_OPT_OVERLAP=1
EXTRA="--skip-cpu-baseline --skip-parity-mode" b ffma_attn FIRA_ATTN_TC=0
EXTRA="--skip-cpu-baseline --skip-parity-mode --layout trimmed" b trimmed X=1
timeout 300 python tools/bench_kernels.py > gpurun_out/r2e_kernels_ab.jsonl 2> gpurun_out/r2e_kernels_ab.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2e_launches.csv \
  python bench.py --steps 2 --warmup 3 --profile-step > gpurun_out/r2e_launches.log 2>&1
for k in attn_tc_fwd attn_tc_bwd gcn_fused_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 3 -f -o gpurun_out/r2e_$k python tools/bench_kernels.py > gpurun_out/r2e_ncu_$k.log 2>&1
done
for f in gpurun_out/r2e_pytest_*.log; do echo "== $f"; tail -n 6 $f; done
head -c 600 gpurun_out/r2e_bench_default.json; echo; tail -n 3 gpurun_out/r2e_bench_default.err
