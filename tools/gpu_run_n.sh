#!/bin/bash
# multi-GPU call (gpurun --gpus N): correctness of the overlapped data-parallel step, bench at N GPUs with and without
# the overlap, config-5 stress sweep on all GPUs.   usage: bash tools/gpu_run_n.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 tools/check_dp_engine.py --layout packed > gpurun_out/r2n${N}_check_packed.json 2> gpurun_out/r2n${N}_check_packed.err
timeout 600 $TR --master-port 29532 tools/check_dp_engine.py --layout padded > gpurun_out/r2n${N}_check_padded.json 2> gpurun_out/r2n${N}_check_padded.err
timeout 600 $TR --master-port 29536 tools/check_dp_engine.py --layout packed --optimizer flat_adam > gpurun_out/r2n${N}_check_flat_adam.json 2> gpurun_out/r2n${N}_check_flat_adam.err
timeout 900 $TR --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2n${N}_bench.json 2> gpurun_out/r2n${N}_bench.err
FIRA_DP_OVERLAP=0 timeout 900 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2n${N}_bench_nooverlap.json 2> gpurun_out/r2n${N}_bench_nooverlap.err
[ -n "$SKIP_STRESS" ] || timeout 900 $TR --master-port 29535 tools/stress_sweep_multi.py > gpurun_out/r2n${N}_stress_sweep.jsonl 2> gpurun_out/r2n${N}_stress_sweep.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2n${N}_bench_n1.json 2> gpurun_out/r2n${N}_bench_n1.err
tail -n 2 gpurun_out/r2n${N}_check_*.json; tail -n 3 gpurun_out/r2n${N}_check_flat_adam.err; for f in bench bench_nooverlap bench_n1; do head -c 300 gpurun_out/r2n${N}_$f.json; echo; done; tail -c 600 gpurun_out/r2n${N}_check_packed.err
