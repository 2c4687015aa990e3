#!/bin/bash
# round 2, call 7: dynamic instruction counts of k_p2g_warp with the two SVD variants + source-level capture of the current kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_7
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:"k_p2g_warp|k_g2p\$|k_g2p_bwd_warp|k_p2g_bwd\$" -s 200 -c 12 --csv --log-file gpurun_out/${T}_inst_new.csv $CMD > gpurun_out/${T}_inst_new.log 2>&1
UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200_hestenes.so timeout 600 ncu --metrics $M --clock-control none -k regex:"k_p2g_warp" -s 100 -c 4 --csv --log-file gpurun_out/${T}_inst_hest.csv $CMD > gpurun_out/${T}_inst_hest.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g_warp\$" -s 150 -c 1 -o gpurun_out/${T}_prof_k_p2g_warp $CMD > gpurun_out/${T}_ncu_p2g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g_bwd\$" -s 40 -c 1 -o gpurun_out/${T}_prof_k_p2g_bwd $CMD > gpurun_out/${T}_ncu_p2gb.log 2>&1
ls -la gpurun_out/${T}_*
