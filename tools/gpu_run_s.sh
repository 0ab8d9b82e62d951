#!/bin/bash
# round-2 GPU call S: vector reductions (red.v4.f32) in the embedding backward and the split-K epilogue; smoke(); whole suite
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2s_smoke.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2s_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2s_bench_$name.json 2> gpurun_out/r2s_bench_$name.err; }
b default X=1
b default2 X=1
timeout 600 python bench.py --steps 10 --warmup 5 --timeline gpurun_out/r2s_timeline.json > gpurun_out/r2s_timeline.log 2>&1
tail -n 6 gpurun_out/r2s_smoke.log; tail -n 6 gpurun_out/r2s_pytest_all.log
python - <<'PY'
import json
for n in ['default','default2']:
    for l in open(f'gpurun_out/r2s_bench_{n}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches']//d['steps'])
PY
