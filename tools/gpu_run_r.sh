#!/bin/bash
# round-2 GPU call R: quarter-warp scatter as the default -- whole suite, variants 8/9/10, ncu of the new default, full bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2r_pytest_all.log
for v in 8 9 10; do FIRA_SPMM_VARIANT=$v timeout 200 python tools/scatter_variants.py >> gpurun_out/r2r_scatter_variants.jsonl 2>> gpurun_out/r2r_scatter_variants.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_spmm_part_kernel -s 8 -c 3 -f -o gpurun_out/r2r_csr_spmm_default python tools/scatter_variants.py > gpurun_out/r2r_ncu_spmm.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2r_bench_full.json 2> gpurun_out/r2r_bench_full.err
for f in gpurun_out/r2r_pytest_*.log; do echo "== $f"; tail -n 8 $f; done
cat gpurun_out/r2r_scatter_variants.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r2r_bench_full.json'):
    if l.startswith('{'):
        d=json.loads(l); print('full', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value']), d['cpu_baseline']['value'])
        for k,v in d.items():
            if k.startswith('roofline') and v: print('  ',k, v.get('rows'), round(v['avg_launch_ms']*1e3,2),'us', round(v['frac'],3))
PY
tail -2 gpurun_out/r2r_bench_full.err
