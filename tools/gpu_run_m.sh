#!/bin/bash
# round-2 GPU call M: relu backward in the dx epilogue, 4-deep bf16 scatter, parallel weight-prep / adjoint groups, gemm_ln fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_ops_bf16.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2m_pytest_tc.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2m_pytest_all.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-parity-mode > gpurun_out/r2m_bench_$name.json 2> gpurun_out/r2m_bench_$name.err; }
b default X=1
b no_dx_relu FIRA_DX_RELU=0
b gemm_ln FIRA_GEMM_LN=1
b default2 X=1
timeout 600 python bench.py --steps 10 --warmup 5 --timeline gpurun_out/r2m_timeline.json > gpurun_out/r2m_timeline.log 2>&1
for f in gpurun_out/r2m_pytest_*.log; do echo "== $f"; tail -n 14 $f; done
python - <<'PY'
import json
for n in ['default','no_dx_relu','gemm_ln','default2']:
    try:
        for l in open(f'gpurun_out/r2m_bench_{n}.json'):
            if l.startswith('{'):
                d=json.loads(l); print(n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches']//d['steps'])
                for k,v in d.items():
                    if k.startswith('roofline') and v and n=='default': print('  ',k, v.get('rows'), round(v['avg_launch_ms']*1e3,2),'us', round(v['frac'],3))
    except Exception as e: print(n,'ERR',e)
PY
tail -3 gpurun_out/r2m_bench_default.err
