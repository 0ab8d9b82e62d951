#!/bin/bash
# round-2 GPU call U: vectorised head kernels; final evidence of the round (smoke, whole suite, full bench line, launch list)
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2u_smoke.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2u_pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench_full.json 2> gpurun_out/r2u_bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2u_launches.csv \
  python bench.py --steps 2 --warmup 3 --profile-step > gpurun_out/r2u_launches.log 2>&1
tail -n 4 gpurun_out/r2u_smoke.log; tail -n 5 gpurun_out/r2u_pytest_all.log
python - <<'PY'
import json
for l in open('gpurun_out/r2u_bench_full.json'):
    if l.startswith('{'):
        d=json.loads(l); print('full', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['e2e_loader']['value']), d['cpu_baseline']['value'], d['fp32_parity_mode']['value'], d['clocks'])
        for k,v in d.items():
            if k.startswith('roofline') and v: print('  ',k, v.get('rows'), round(v['avg_launch_ms']*1e3,2),'us', round(v['frac'],3))
PY
grep -c "head_fwd\|head_bwd" gpurun_out/r2u_launches.csv; grep "head_fwd_kernel\|head_bwd_kernel" gpurun_out/r2u_launches.csv | grep duration | cut -c1-60,200-400 | head -4
