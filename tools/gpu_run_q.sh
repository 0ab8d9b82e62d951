#!/bin/bash
# round-2 GPU call Q: evidence of the final build -- scatter variants, launch list + ncu --set full captures of one step's
# kernels, the full default bench line and the reference arm
mkdir -p gpurun_out
for v in 4 1 6 7 8; do FIRA_SPMM_VARIANT=$v timeout 200 python tools/scatter_variants.py >> gpurun_out/r2q_scatter_variants.jsonl 2>> gpurun_out/r2q_scatter_variants.err; done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench_full.json 2> gpurun_out/r2q_bench_full.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2q_launches.csv \
  python bench.py --steps 2 --warmup 3 --profile-step > gpurun_out/r2q_launches.log 2>&1
for k in gemm_tc_kernel csr_spmm_part_kernel adam_flat_kernel attn_tc_fwd_kernel attn_tc_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 4 -f -o gpurun_out/r2q_$k \
    python bench.py --steps 2 --warmup 3 --profile-step > gpurun_out/r2q_ncu_$k.log 2>&1
done
cat gpurun_out/r2q_scatter_variants.jsonl
python - <<'PY'
import json
for n in ['full','reference']:
    for l in open(f'gpurun_out/r2q_bench_{n}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(n, round(d['value'],1), round(d['ms_per_step'],3), d.get('e2e'), d.get('cpu_baseline',{}) and d['cpu_baseline'].get('value'))
PY
tail -2 gpurun_out/r2q_bench_full.err; tail -2 gpurun_out/r2q_launches.log; ls -la gpurun_out/r2q_*.ncu-rep
