#!/bin/bash
# Round-2 FINAL measurements (one gpurun call) of the build with the shell-job skip, the specialised grid kernels, the
# alternating raw / cotangent grids and the parallel FK adjoint.  Everything lands in gpurun_out/ and is summarised into
# profiles/ by profiles/summarize.py afterwards.  A command runs under ncu only after it exited 0 without it.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02g
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v Warning > gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
for c in pour_water whip_rope cloth_para; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c >> gpurun_out/${T}_other_configs.jsonl 2>> gpurun_out/${T}_other.err
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --config whip_rope --ckpt-window 70 >> gpurun_out/${T}_other_configs.jsonl 2>> gpurun_out/${T}_other.err
for t in "--env-groups 1" "--env-groups 3" "--env-groups 4" "--env-groups 1 --tune pers=0" "--p2g-mode 1" "--adjoint tape"; do
  python bench.py --no-cpu-baseline --no-e2e --steps 10 $t > "gpurun_out/${T}_ab_$(echo $t | tr ' =' '__').json" 2>/dev/null
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --env-groups 1"
$CMD > gpurun_out/${T}_profiled_cmd.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1
for k in k_p2g_pers k_g2p; do
  ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" -s 150 -c 1 -o gpurun_out/${T}_prof_$k $CMD > gpurun_out/${T}_ncu_$k.log 2>&1
done
for k in k_p2g_bwd k_g2p_bwd_warp k_grid_bwd k_grid_fwd; do
  ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" -s 40 -c 1 -o gpurun_out/${T}_prof_$k $CMD > gpurun_out/${T}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${T}_*
