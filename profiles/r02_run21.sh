#!/bin/bash
# round 2, calls 21-22: node-tile stride 33 vs 32 float4 per cell (bank conflicts of the 27 stencil loads).  Call 21 also
# carried a dynamic tile queue (atomic tickets) for the persistent P2G / P2G^T: 126 -> 139 and 142 -> 155 us, removed.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_21
run() {  # name lib tune
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --env-groups 1 $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_21_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p')})
PY
}
for i in 1 2; do
  run product "" ""
  run nt32 _nt32 ""
done
timeout 900 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
