#!/bin/bash
# round 2, call 47: grid update of multi-primitive scenes skips, per warp, a primitive none of its cells is within reach
# of (pour_water: two bowls).  A/B against the previous build + the tests that cover pour_water and two-primitive scenes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_44
run() {  # name lib args
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_44_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p','grid','grid_bwd')})
PY
}
run base_pour _base "--config pour_water"
run new_pour "" "--config pour_water"
run base_push _base "--env-groups 1"
run new_push "" "--env-groups 1"
timeout 600 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_mpmenv_gpu.py tests/test_fullsize_parity_gpu.py -q -s -m gpu -k "not shape_rope_env" -o faulthandler_timeout=200 > gpurun_out/${T}_tests.log 2>&1; tail -2 gpurun_out/${T}_tests.log
