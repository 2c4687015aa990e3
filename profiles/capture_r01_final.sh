# Round-1 final measurements (one gpurun call).  Everything lands in gpurun_out/ and is summarised into profiles/ by
# profiles/summarize.py / kern_table.py afterwards.  A command runs under ncu only after it exited 0 without it.
set -x
T=r01f
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
for t in sort=0 stage=0 sort=0,stage=0 svd_warm=0; do
  python bench.py --no-cpu-baseline --no-e2e --steps 10 --tune $t > gpurun_out/${T}_ab_$t.json 2>/dev/null
done
python bench.py --no-cpu-baseline --no-e2e --steps 10 --p2g-mode 1 > gpurun_out/${T}_ab_det.json 2>/dev/null
python profiles/other_configs.py > gpurun_out/${T}_other_configs.jsonl 2> gpurun_out/${T}_other.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${T}_profiled_cmd.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1
for k in k_p2g k_g2p; do
  ncu --set full --clock-control none --import-source on -k regex:"${k}\$" -s 150 -c 1 -o gpurun_out/${T}_prof_$k $CMD > gpurun_out/${T}_ncu_$k.log 2>&1
done
for k in k_p2g_bwd k_g2p_bwd k_grid_bwd k_grid_fwd; do
  ncu --set full --clock-control none --import-source on -k regex:"${k}\$" -s 40 -c 1 -o gpurun_out/${T}_prof_$k $CMD > gpurun_out/${T}_ncu_$k.log 2>&1
done
for t in sort=0 stage=0; do
  ncu --set full --clock-control none -k regex:"k_p2g\$" -s 150 -c 1 -o gpurun_out/${T}_prof_p2g_$t $CMD --tune $t > gpurun_out/${T}_ncu_p2g_$t.log 2>&1
done
ls -la gpurun_out/${T}_*
