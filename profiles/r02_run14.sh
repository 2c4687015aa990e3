#!/bin/bash
# round 2, call 14: coalesced un-sort stores, bare-weight staging, k_g2p at 8 / 6 / 5 CTAs per SM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_14
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_fullsize_parity_gpu.py 2>&1 | grep -v Warning | tail -12 > gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
for v in "" _g2p6 _g2p5; do
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$v.so timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench$v.json 2> gpurun_out/${T}_bench$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_14_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
