#!/bin/bash
# round-2 GPU call H: optim.FlatAdam (flat parameters, in-place gradients, bf16 mirror), side streams, PDL A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2h_pytest_optim.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2h_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2h_bench_$name.json 2> gpurun_out/r2h_bench_$name.err; }
b default X=1
b pdl FIRA_PDL=1
b pdl_side8 FIRA_PDL=1 FIRA_SIDE_STREAMS=8
b torch_adam FIRA_TORCH_ADAM=1
b pdl_fused FIRA_PDL=1 FIRA_GCN_FUSED=1
FIRA_PDL=1 timeout 600 python bench.py --steps 10 --warmup 5 --timeline gpurun_out/r2h_timeline_pdl.json > gpurun_out/r2h_timeline.log 2>&1
for f in gpurun_out/r2h_pytest_*.log; do echo "== $f"; tail -n 12 $f; done
for n in default pdl pdl_side8 torch_adam pdl_fused; do head -c 160 gpurun_out/r2h_bench_$n.json | cut -c40-160; echo; tail -n 3 gpurun_out/r2h_bench_$n.err; done
