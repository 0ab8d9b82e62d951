#!/bin/bash
# 8-GPU call (gpurun --gpus 8): overlapped data-parallel step checked on 8 ranks, bench at N = 8, config-5 stress sweep on
# all 8 GPUs.  Charged 8x: only what needs 8 GPUs runs here.
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29536 tools/check_dp_engine.py --layout packed --optimizer flat_adam > gpurun_out/r2n8_check_flat_adam.json 2> gpurun_out/r2n8_check_flat_adam.err
timeout 500 $TR --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err
timeout 500 $TR --master-port 29535 tools/stress_sweep_multi.py > gpurun_out/r2n8_stress_sweep.jsonl 2> gpurun_out/r2n8_stress_sweep.err
grep "^{" gpurun_out/r2n8_check_flat_adam.json; grep "^{" gpurun_out/r2n8_bench.json | head -c 400; echo; tail -n 4 gpurun_out/r2n8_stress_sweep.jsonl; tail -n 3 gpurun_out/r2n8_bench.err | cut -c1-300
