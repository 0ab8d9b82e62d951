#!/bin/bash
# round-2 GPU call B: fixed/new tests in their own processes, single-kernel A/B timings, launch list of the step with the
# new tcgen05 kernels, bench variants (fused GCN / tcgen05 attention / packed layout)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_blocks.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2b_pytest_blocks.log
timeout 600 python -m pytest tests/test_gpu_zzzz_gcn_fused.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2b_pytest_gcn_fused.log
timeout 900 python -m pytest tests/test_gpu_packed.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r2b_pytest_packed.log
FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1 timeout 900 python -m pytest tests/test_gpu_packed.py -m gpu -q -s -k bf16 2>&1 | tail -60 > gpurun_out/r2b_pytest_packed_newkernels.log
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r2b_pytest_engine.log
timeout 600 python tools/bench_kernels.py > gpurun_out/r2b_kernels.jsonl 2> gpurun_out/r2b_kernels.err
FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 2600 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --no-graph > gpurun_out/r2b_ncu_bench.log 2>&1
FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1 timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2b_bench_bf16_both.json 2> gpurun_out/r2b_bench_bf16_both.err
timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --layout packed > gpurun_out/r2b_bench_bf16_packed.json 2> gpurun_out/r2b_bench_bf16_packed.err
FIRA_ATTN_TC=1 FIRA_GCN_FUSED=1 timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --layout packed > gpurun_out/r2b_bench_bf16_packed_both.json 2> gpurun_out/r2b_bench_bf16_packed_both.err
tail -3 gpurun_out/r2b_pytest_blocks.log gpurun_out/r2b_pytest_gcn_fused.log gpurun_out/r2b_pytest_packed.log gpurun_out/r2b_pytest_packed_newkernels.log gpurun_out/r2b_pytest_engine.log
