#!/bin/bash
# round-2 GPU call A: full gpu test-suite (no -x), new tcgen05 kernels (fused GCN, attention) in their own processes,
# pipelined scatter variant, bench lines (bf16 default / with the new kernels, fp32, reference arm)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --ignore=tests/test_gpu_zzzz_gcn_fused.py --ignore=tests/test_gpu_packed.py 2>&1 | tail -80 > gpurun_out/r2a_pytest.log
timeout 900 python -m pytest tests/test_gpu_packed.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2a_pytest_packed.log
FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1 timeout 900 python -m pytest tests/test_gpu_packed.py -m gpu -q -s -k bf16 2>&1 | tail -60 > gpurun_out/r2a_pytest_packed_newkernels.log
timeout 600 python -m pytest tests/test_gpu_zzzz_gcn_fused.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r2a_pytest_gcn_fused.log
FIRA_ATTN_TC=1 timeout 600 python -m pytest tests/test_gpu_ops_bf16.py -m gpu -q -s -k attention 2>&1 | tail -80 > gpurun_out/r2a_pytest_attn_tc.log
FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_engine.py -m gpu -q -s -k "bf16 or graph" 2>&1 | tail -60 > gpurun_out/r2a_pytest_model_newkernels.log
( FIRA_SPMM_VARIANT=6 timeout 300 python tools/check_spmm_variant.py --time ) > gpurun_out/r2a_spmm_v6.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_bf16.json 2> gpurun_out/r2a_bench_bf16.err
FIRA_GCN_FUSED=1 timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2a_bench_bf16_fused.json 2> gpurun_out/r2a_bench_bf16_fused.err
FIRA_ATTN_TC=1 timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2a_bench_bf16_attntc.json 2> gpurun_out/r2a_bench_bf16_attntc.err
FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1 timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2a_bench_bf16_both.json 2> gpurun_out/r2a_bench_bf16_both.err
timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --layout packed > gpurun_out/r2a_bench_bf16_packed.json 2> gpurun_out/r2a_bench_bf16_packed.err
FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1 timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --layout packed > gpurun_out/r2a_bench_bf16_packed_both.json 2> gpurun_out/r2a_bench_bf16_packed_both.err
timeout 600 python bench.py --steps 20 --warmup 5 --precision fp32 --skip-cpu-baseline > gpurun_out/r2a_bench_fp32.json 2> gpurun_out/r2a_bench_fp32.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
tail -5 gpurun_out/r2a_pytest.log; tail -5 gpurun_out/r2a_pytest_gcn_fused.log; tail -5 gpurun_out/r2a_pytest_attn_tc.log; tail -5 gpurun_out/r2a_pytest_packed.log
