#!/bin/bash
# round-2 GPU call F: programmatic dependent launch (PDL) on every kernel -- whole suite, A/B bench lines, kernel timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2f_pytest_all.log
FIRA_OPT_OVERLAP=1 timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_train_curve.py -m gpu -q 2>&1 | tail -20 > gpurun_out/r2f_pytest_overlap.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2f_bench_$name.json 2> gpurun_out/r2f_bench_$name.err; }
b pdl X=1
b nopdl FIRA_PDL=0
b pdl_overlap FIRA_OPT_OVERLAP=1
timeout 600 python bench.py --steps 10 --warmup 5 --timeline gpurun_out/r2f_timeline_pdl.json > gpurun_out/r2f_timeline_pdl.log 2>&1
FIRA_PDL=0 timeout 600 python bench.py --steps 10 --warmup 5 --timeline gpurun_out/r2f_timeline_nopdl.json > gpurun_out/r2f_timeline_nopdl.log 2>&1
for f in gpurun_out/r2f_pytest_*.log; do echo "== $f"; tail -n 5 $f; done
for n in pdl nopdl pdl_overlap; do head -c 200 gpurun_out/r2f_bench_$n.json; echo; tail -n 2 gpurun_out/r2f_bench_$n.err; done
