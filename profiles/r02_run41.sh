#!/bin/bash
# round 2, call 44: node tiles of k_g2p / k_p2g_bwd for up to 8 distinct cells per warp (-DUD_G2P_TILE_CELLS=8) instead
# of 4: warps of low-density scenes (pour_water: ~8 cells per warp) otherwise gather their 27 nodes from L1/L2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_41
run() {  # name lib args
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_41_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p','grid','grid_bwd')})
PY
}
run c4_pour "" "--config pour_water"
run c8_pour _c8 "--config pour_water"
run c4_whip "" "--config whip_rope"
run c8_whip _c8 "--config whip_rope"
run c4_push "" "--env-groups 1"
run c8_push _c8 "--env-groups 1"
run c8_push_g2 _c8 ""
