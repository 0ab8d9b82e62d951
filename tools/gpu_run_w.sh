#!/bin/bash
# round-2 GPU call W (last minutes of the budget): copy_scores_fwd around the active target rows -- whole suite + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2w_pytest_all.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
tail -n 4 gpurun_out/r2w_pytest_all.log; grep "^{" gpurun_out/r2w_bench.json | head -c 260; echo; tail -2 gpurun_out/r2w_bench.err
