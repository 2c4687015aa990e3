#!/bin/bash
# Round-2 FINAL measurements (one gpurun call) of the final build: liquid path behind UD_P2G_LIQUID_FAST, crowded-cell
# ranking (k_rank_big), fused APG update in its two forms.  Everything lands in gpurun_out/ and is summarised into
# profiles/ by profiles/summarize.py afterwards.  A command runs under ncu only after it exited 0 without it.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02h
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -o faulthandler_timeout=300 2>&1 | grep -v Warning > gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
for c in pour_water whip_rope cloth_para; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c >> gpurun_out/${T}_other_configs.jsonl 2>> gpurun_out/${T}_other.err
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --config whip_rope --ckpt-window 70 >> gpurun_out/${T}_other_configs.jsonl 2>> gpurun_out/${T}_other.err
for t in "--env-groups 1" "--p2g-mode 1" "--adjoint tape"; do
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
# the liquid kernels on the pour_water scene
CMDP="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --env-groups 1 --config pour_water"
ncu --set full --clock-control none --import-source on -k regex:"^k_p2g_pers\$" -s 150 -c 1 -o gpurun_out/${T}_prof_pour_k_p2g_pers $CMDP > gpurun_out/${T}_ncu_pour_p2g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_p2g_bwd\$" -s 30 -c 1 -o gpurun_out/${T}_prof_pour_k_p2g_bwd $CMDP > gpurun_out/${T}_ncu_pour_p2g_bwd.log 2>&1
ls -la gpurun_out/${T}_*
