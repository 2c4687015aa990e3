set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_r01z.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01z.csv $CMD > gpurun_out/ncu_r01z_launch.log 2>&1
for k in k_p2g k_g2p; do
  ncu --set full --clock-control none --import-source on -k regex:"${k}\$" -s 150 -c 1 -o gpurun_out/prof_r01z_$k $CMD > gpurun_out/ncu_r01z_$k.log 2>&1
done
for k in k_p2g_bwd k_g2p_bwd k_grid_bwd; do
  ncu --set full --clock-control none --import-source on -k regex:"${k}\$" -s 40 -c 1 -o gpurun_out/prof_r01z_$k $CMD > gpurun_out/ncu_r01z_$k.log 2>&1
done
for t in sort=0 stage=0; do
  ncu --set full --clock-control none -k regex:"k_p2g\$" -s 150 -c 1 -o gpurun_out/prof_r01z_p2g_$t $CMD --tune $t > gpurun_out/ncu_r01z_p2g_$t.log 2>&1
done
ls -la gpurun_out/*r01z*
