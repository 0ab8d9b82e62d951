#!/bin/bash
# round-2 GPU call D: tcgen05 attention on by default, Adam overlap re-test, full suite in one process (what the driver
# runs), bench variants, ncu --set full captures of the new tensor-core kernels
mkdir -p gpurun_out
FIRA_OPT_OVERLAP=1 timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_train_curve.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2d_pytest_overlap.log
timeout 1400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2d_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline $EXTRA > gpurun_out/r2d_bench_$name.json 2> gpurun_out/r2d_bench_$name.err; }
EXTRA="" b default X=1
EXTRA="" b overlap FIRA_OPT_OVERLAP=1
EXTRA="" b fused FIRA_GCN_FUSED=1
EXTRA="" b fused_overlap FIRA_GCN_FUSED=1 FIRA_OPT_OVERLAP=1
EXTRA="--layout trimmed" b trimmed X=1
EXTRA="--layout trimmed" b trimmed_fused FIRA_GCN_FUSED=1
for k in attn_tc_fwd attn_tc_bwd gcn_fused_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 3 -f -o gpurun_out/r2d_$k python tools/bench_kernels.py > gpurun_out/r2d_ncu_$k.log 2>&1
done
for f in gpurun_out/r2d_pytest_*.log; do echo "== $f"; tail -n 4 $f; done
