#!/bin/bash
# round 2, call 23: one-wave-ahead L2 prefetch in the one-tile-per-warp kernels (G2P, G2P^T); packed fp32 staging in P2G
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_22
run() {  # name lib tune
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --env-groups 1 $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_22_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p')})
PY
}
for i in 1 2; do
  run product "" ""
  run stagescalar _stagescalar ""
  run noahead _noahead ""
  run ahead2 _ahead2 ""
done
timeout 900 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
