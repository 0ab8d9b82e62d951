#!/bin/bash
# round-2 GPU call T: vector reductions in the LayerNorm backward; beam-search numbers of the final build (config 4)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_ops_bf16.py tests/test_gpu_model.py tests/test_gpu_train_curve.py tests/test_gpu_cli.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r2t_pytest.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2t_bench_$name.json 2> gpurun_out/r2t_bench_$name.err; }
b default X=1
b default2 X=1
timeout 600 python tools/bench_beam.py --batches 20,128 --beams 3,5 --precision fp32 > gpurun_out/r2t_beam_fp32.jsonl 2> gpurun_out/r2t_beam_fp32.err
timeout 600 python tools/bench_beam.py --batches 20,128 --beams 3,5 --precision bf16 > gpurun_out/r2t_beam_bf16.jsonl 2> gpurun_out/r2t_beam_bf16.err
tail -n 5 gpurun_out/r2t_pytest.log
python - <<'PY'
import json
for n in ['default','default2']:
    for l in open(f'gpurun_out/r2t_bench_{n}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))
for f in ['fp32','bf16']:
    for l in open(f'gpurun_out/r2t_beam_{f}.jsonl'):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['batch'], d['beam'], d['mode'], round(d['value'],1), d['ids_equal_full_mode'])
PY
tail -2 gpurun_out/r2t_beam_fp32.err
