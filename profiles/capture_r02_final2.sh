#!/bin/bash
# Round-2 final capture, second part (after the multi-primitive skip in k_grid_fwd, the only kernel that changed since
# capture_r02_final.sh): smoke, full GPU suite, default bench, reference arm, the other configs, launch list, k_grid_fwd
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02i
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
timeout 1000 python -m pytest tests -m gpu -q -s -o faulthandler_timeout=300 2>&1 | grep -v Warning > gpurun_out/${T}_tests.log
tail -2 gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
for c in pour_water whip_rope cloth_para; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c >> gpurun_out/${T}_other_configs.jsonl 2>> gpurun_out/${T}_other.err
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --config whip_rope --ckpt-window 70 >> gpurun_out/${T}_other_configs.jsonl 2>> gpurun_out/${T}_other.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --env-groups 1"
$CMD > gpurun_out/${T}_profiled_cmd.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_grid_fwd\$" -s 40 -c 1 -o gpurun_out/${T}_prof_k_grid_fwd $CMD > gpurun_out/${T}_ncu_k_grid_fwd.log 2>&1
CMDP="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --env-groups 1 --config pour_water"
ncu --set full --clock-control none --import-source on -k regex:"^k_grid_fwd\$" -s 60 -c 1 -o gpurun_out/${T}_prof_pour_k_grid_fwd $CMDP > gpurun_out/${T}_ncu_pour_grid.log 2>&1
ls gpurun_out/${T}_* | wc -l
