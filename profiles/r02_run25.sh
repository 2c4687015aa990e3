#!/bin/bash
# round 2, call 26: one-tile-per-warp kernels (G2P, G2P^T) walking the tiles downwards after an upward persistent kernel
# (L2 reuse of what that kernel wrote last); ncu source-level captures of the five hot kernels at HEAD
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_25
run() {  # name lib args
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_25_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p','grid','grid_bwd')})
PY
}
for i in 1 2; do
  run base$i "" "--env-groups 1"
  run rev$i _rev "--env-groups 1"
done
run base_g2 "" ""
run rev_g2 _rev ""
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --env-groups 1"
for k in k_p2g_pers k_g2p; do
  ncu --set full --clock-control none --import-source on -k regex:"${k}\$" -s 150 -c 1 -o gpurun_out/${T}_prof_$k $CMD > gpurun_out/${T}_ncu_$k.log 2>&1
done
for k in k_p2g_bwd k_g2p_bwd_warp k_grid_bwd; do
  ncu --set full --clock-control none --import-source on -k regex:"${k}\$" -s 40 -c 1 -o gpurun_out/${T}_prof_$k $CMD > gpurun_out/${T}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${T}_*
