#!/bin/bash
# round 2, call 9: all four particle kernels persistent with next-tile prefetch; k_g2p at 8 vs 6 CTAs per SM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_9
timeout 900 python -m pytest tests/test_mpm_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | grep -v Warning | tail -30 > gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
for v in "" _g2p6; do
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$v.so timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench$v.json 2> gpurun_out/${T}_bench$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_9_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
