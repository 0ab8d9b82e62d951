#!/bin/bash
# round-2 GPU call A: full gpu test-suite (no -x), fused-GCN kernel tests in their own process, pipelined scatter variant,
# bench lines (bf16 default, fp32, reference arm)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --ignore=tests/test_gpu_zzzz_gcn_fused.py 2>&1 | tail -80 > gpurun_out/r2a_pytest.log
timeout 600 python -m pytest tests/test_gpu_zzzz_gcn_fused.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r2a_pytest_gcn_fused.log
( FIRA_SPMM_VARIANT=6 timeout 300 python tools/check_spmm_variant.py --time ) > gpurun_out/r2a_spmm_v6.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_bf16.json 2> gpurun_out/r2a_bench_bf16.err
FIRA_GCN_FUSED=1 timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2a_bench_bf16_fused.json 2> gpurun_out/r2a_bench_bf16_fused.err
timeout 600 python bench.py --steps 20 --warmup 5 --precision fp32 --skip-cpu-baseline > gpurun_out/r2a_bench_fp32.json 2> gpurun_out/r2a_bench_fp32.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
tail -5 gpurun_out/r2a_pytest.log; tail -5 gpurun_out/r2a_pytest_gcn_fused.log
