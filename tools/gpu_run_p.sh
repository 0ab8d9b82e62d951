#!/bin/bash
# round-2 GPU call P: side-work tail (lane-0 group split), wgrad split-K target A/B; whole suite
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2p_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2p_bench_$name.json 2> gpurun_out/r2p_bench_$name.err; }
b default X=1
b wgrad32 FIRA_WGRAD_CTAS=32
b wgrad64 FIRA_WGRAD_CTAS=64
b wgrad16 FIRA_WGRAD_CTAS=16
b default2 X=1
for f in gpurun_out/r2p_pytest_*.log; do echo "== $f"; tail -n 8 $f; done
python - <<'PY'
import json
for n in ['default','wgrad32','wgrad64','wgrad16','default2']:
    try:
        for l in open(f'gpurun_out/r2p_bench_{n}.json'):
            if l.startswith('{'):
                d=json.loads(l); print(n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches']//d['steps'])
    except Exception as e: print(n,'ERR',e)
PY
tail -3 gpurun_out/r2p_bench_default.err
