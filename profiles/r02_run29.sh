#!/bin/bash
# round 2, call 30: ncu launch lists of the low-density / dense-cell configs (whip_rope: sort class 183 us per launch;
# pour_water: grid / grid_bwd with two container colliders)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_29
for c in whip_rope pour_water; do
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --env-groups 1 --config $c"
  $CMD > gpurun_out/${T}_$c.json 2> gpurun_out/${T}_$c.err || exit 1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${T}_launches_$c.csv $CMD > gpurun_out/${T}_ncu_$c.log 2>&1
done
ls -la gpurun_out/${T}_*
