set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_r01t.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_p2g\$" -s 150 -c 1 -o gpurun_out/prof_r01t_k_p2g $CMD > gpurun_out/ncu_r01t_k_p2g.log 2>&1
for k in k_p2g_bwd k_g2p_bwd; do
  ncu --set full --clock-control none --import-source on -k regex:"${k}\$" -s 40 -c 1 -o gpurun_out/prof_r01t_$k $CMD > gpurun_out/ncu_r01t_$k.log 2>&1
done
ls -la gpurun_out/*r01t*
