#!/bin/bash
# round-2 GPU call V: final validation of the committed build (whole suite, full bench line)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2v_pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench_full.json 2> gpurun_out/r2v_bench_full.err
tail -n 5 gpurun_out/r2v_pytest_all.log
python - <<'PY'
import json
for l in open('gpurun_out/r2v_bench_full.json'):
    if l.startswith('{'):
        d=json.loads(l); print('full', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['e2e_loader']['value']), d['cpu_baseline']['value'], d['fp32_parity_mode']['value'], d['clocks'])
PY
tail -2 gpurun_out/r2v_bench_full.err
