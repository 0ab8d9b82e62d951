#!/bin/bash
# round-2 GPU call O: backward side work joined once at the end of the backward pass (and into the optimizer stream)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2o_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2o_bench_$name.json 2> gpurun_out/r2o_bench_$name.err; }
b default X=1
b no_defer FIRA_DEFER_JOIN=0
b default2 X=1
b no_defer2 FIRA_DEFER_JOIN=0
timeout 600 python bench.py --steps 10 --warmup 5 --timeline gpurun_out/r2o_timeline.json > gpurun_out/r2o_timeline.log 2>&1
for f in gpurun_out/r2o_pytest_*.log; do echo "== $f"; tail -n 14 $f; done
python - <<'PY'
import json
for n in ['default','no_defer','default2','no_defer2']:
    try:
        for l in open(f'gpurun_out/r2o_bench_{n}.json'):
            if l.startswith('{'):
                d=json.loads(l); print(n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches']//d['steps'])
    except Exception as e: print(n,'ERR',e)
PY
tail -3 gpurun_out/r2o_bench_default.err
