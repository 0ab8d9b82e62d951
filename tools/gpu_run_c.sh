#!/bin/bash
# round-2 GPU call C: v2 tcgen05 attention, GEMM tile-width heuristic, 16-bit dropout RNG, Adam overlap, packed default
mkdir -p gpurun_out
FIRA_ATTN_TC=1 timeout 600 python -m pytest tests/test_gpu_ops_bf16.py -m gpu -q -k attention 2>&1 | tail -40 > gpurun_out/r2c_pytest_attn_tc.log
FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1 timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_model.py tests/test_gpu_zzzz_gcn_fused.py -m gpu -q -k "bf16 or fused or gcn_layer" 2>&1 | tail -40 > gpurun_out/r2c_pytest_newkernels.log
timeout 1200 python -m pytest tests/test_gpu_train_curve.py tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_cli.py tests/test_gpu_ops.py tests/test_gpu_ops_bf16.py tests/test_gpu_packed.py -m gpu -q 2>&1 | tail -60 > gpurun_out/r2c_pytest_main.log
timeout 600 python tools/bench_kernels.py > gpurun_out/r2c_kernels.jsonl 2> gpurun_out/r2c_kernels.err
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline $EXTRA > gpurun_out/r2c_bench_$name.json 2> gpurun_out/r2c_bench_$name.err; }
EXTRA="" b packed_default X=1
EXTRA="" b packed_noadamoverlap FIRA_OPT_OVERLAP=0
EXTRA="" b packed_attntc FIRA_ATTN_TC=1
EXTRA="" b packed_fused FIRA_GCN_FUSED=1
EXTRA="" b packed_both FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1
EXTRA="--layout trimmed" b trimmed_default X=1
EXTRA="--layout trimmed" b trimmed_both FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1
for f in gpurun_out/r2c_pytest_*.log; do echo "== $f"; tail -n 3 $f; done
