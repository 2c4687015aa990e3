#!/bin/bash
# round 2, call 31: liquid particles without an SVD (J = |det F1|, dF1 = gJ sign cof F1) in P2G / P2G^T; crowded cells
# ranked by k_rank_big (bitmap + popcount scan).  A/B against the previous build on the three MPM configs + MPM test set.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_30
run() {  # name lib args
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_30_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), d['peak_hbm_bytes'], {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p','grid','grid_bwd','fk','sort','gather','memset','unsort','finish_bwd')})
PY
}
timeout 1200 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_fullsize_gpu.py tests/test_fullsize_parity_gpu.py tests/test_mpmenv_gpu.py -q -s -m gpu -k "not shape_rope_env" > gpurun_out/${T}_tests.log 2>&1; tail -5 gpurun_out/${T}_tests.log
run base_push _base "--env-groups 1"
run new_push "" "--env-groups 1"
run base_pour _base "--config pour_water"
run new_pour "" "--config pour_water"
run base_whip _base "--config whip_rope"
run new_whip "" "--config whip_rope"
run new_push_g2 "" ""
