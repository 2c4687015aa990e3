#!/bin/bash
# round 2, call 24: tile walk by additions instead of a division per tile; interior cells of the node tiles as one
# shuffle + linear index; bulk L2 prefetch (one instruction per tile range) vs per-lane prefetches; SVD-VJP divisions 6 -> 3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_23
run() {  # name lib tune
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --env-groups 1 $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_23_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p')})
PY
}
for i in 1 2; do
  run base _base ""
  run product "" ""
  run lanepf _lanepf ""
done
timeout 900 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_fullsize_gpu.py tests/test_fullsize_parity_gpu.py tests/test_mpmenv_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
