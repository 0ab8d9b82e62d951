#!/bin/bash
# round-2 GPU call I: in-kernel phase probe of the small tcgen05 GEMMs, FlatAdam tests again (tolerance fix), default bench
mkdir -p gpurun_out
timeout 300 python tools/gemm_probe.py > gpurun_out/r2i_gemm_probe.jsonl 2> gpurun_out/r2i_gemm_probe.err
timeout 600 python -m pytest tests/test_gpu_optim.py tests/test_gpu_engine.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_default.err
cat gpurun_out/r2i_gemm_probe.jsonl; tail -3 gpurun_out/r2i_gemm_probe.err; tail -4 gpurun_out/r2i_pytest.log; head -c 200 gpurun_out/r2i_bench_default.json
