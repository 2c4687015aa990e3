#!/bin/bash
# round 2, call 8: persistent P2G with next-tile prefetch (pers=1) vs one tile per warp (pers=0); single-instruction rsqrt/div/sqrt
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_8
timeout 900 python -m pytest tests/test_mpm_gpu.py tests/test_golden_gpu.py -m gpu -q 2>&1 | grep -v Warning | tail -30 > gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
for t in pers=1 pers=0; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --tune $t > gpurun_out/${T}_bench_$t.json 2> gpurun_out/${T}_bench_$t.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_8_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
