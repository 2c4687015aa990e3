#!/bin/bash
# round 2, call 20: packed fp32 adds (FADD2) in the flush's column sums; plastic stress in the SVD frame in the forward P2G
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_20
for v in _base _flushonly "" _base _flushonly ""; do
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$v.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --env-groups 1 > gpurun_out/${T}_bench$v.json 2> gpurun_out/${T}_bench$v.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_20_bench{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1] or 'product', round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p')})
PY
done
timeout 900 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_fullsize_parity_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
