#!/bin/bash
# round-2 GPU call L: fused Linear + dropout + residual + LayerNorm (fira_gemm_ln_fwd)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2l_pytest_tc.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2l_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2l_bench_$name.json 2> gpurun_out/r2l_bench_$name.err; }
b default X=1
b no_gemm_ln FIRA_GEMM_LN=0
# (run L also timed the step with every side-stream launch dropped through a temporary hook, removed since: 2.58 ms)
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench_full.json 2> gpurun_out/r2l_bench_full.err
for f in gpurun_out/r2l_pytest_*.log; do echo "== $f"; tail -n 14 $f; done
python - <<'PY'
import json
for n in ['default','no_gemm_ln','full']:
    try:
        for l in open(f'gpurun_out/r2l_bench_{n}.json'):
            if l.startswith('{'):
                d=json.loads(l); print(n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches']//d['steps'])
    except Exception as e: print(n,'ERR',e)
PY
tail -3 gpurun_out/r2l_bench_default.err; tail -3 gpurun_out/r2l_bench_full.err
python - <<'PY'
import json
for l in open('gpurun_out/r2l_bench_full.json'):
    if l.startswith('{'):
        d=json.loads(l)
        for k,v in d.items():
            if k.startswith('roofline') and v: print(k, v.get('rows'), round(v['avg_launch_ms']*1e3,2),'us', round(v['frac'],3), v.get('frac_tensor'))
        print('cpu', d.get('cpu_baseline')); print('fp32', d.get('fp32_parity_mode'))
PY
