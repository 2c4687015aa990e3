#!/bin/bash
# round 2, call 3: source-level ncu captures of the warp-local kernels (fuse=0) to see where the instructions go
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_3
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/${T}_cmd.json 2> gpurun_out/${T}_cmd.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1
for k in k_p2g_warp k_g2p; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${k}\$" -s 150 -c 1 -o gpurun_out/${T}_prof_$k $CMD > gpurun_out/${T}_ncu_$k.log 2>&1
done
for k in k_p2g_bwd k_g2p_bwd_warp k_grid_bwd k_grid_fwd; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${k}\$" -s 40 -c 1 -o gpurun_out/${T}_prof_$k $CMD > gpurun_out/${T}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${T}_*
