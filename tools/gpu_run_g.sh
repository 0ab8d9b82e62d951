#!/bin/bash
# round-2 GPU call G: side work rotated over several streams, split-K weight merges, optimizer overlap on by default
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2g_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2g_bench_$name.json 2> gpurun_out/r2g_bench_$name.err; }
b default X=1
b side1 FIRA_SIDE_STREAMS=1
b side2 FIRA_SIDE_STREAMS=2
b side8 FIRA_SIDE_STREAMS=8
b nosplit FIRA_MERGE_SPLITS=1
b noopt FIRA_OPT_OVERLAP=0
b pdl FIRA_PDL=1
timeout 600 python bench.py --steps 10 --warmup 5 --timeline gpurun_out/r2g_timeline.json > gpurun_out/r2g_timeline.log 2>&1
for f in gpurun_out/r2g_pytest_*.log; do echo "== $f"; tail -n 5 $f; done
for n in default side1 side2 side8 nosplit noopt pdl; do head -c 160 gpurun_out/r2g_bench_$n.json | cut -c40-160; echo; tail -n 2 gpurun_out/r2g_bench_$n.err; done
