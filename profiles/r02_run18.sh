#!/bin/bash
# round 2, call 18: flush row loop, 8 rows / two accumulators vs 4 rows / one accumulator (segment table in both)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_18
for v in "" _flush4 "" _flush4; do
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$v.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --env-groups 1 > gpurun_out/${T}_bench$v.json 2> gpurun_out/${T}_bench$v.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_18_bench{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1] or 'flush8', round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd')})
PY
done
