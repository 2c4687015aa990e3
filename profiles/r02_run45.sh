#!/bin/bash
# round 2, call 49: the north_star's design-choice A/B runs on the FINAL build: without the per-frame sort, without the
# shared-memory staging (every particle issues its 27 vector REDs), without both, and the round-1 CTA-staged kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_45
for t in "svd_warm=0" "sort=0" "stage=0" "sort=0,stage=0" "warp=0" "pers=0"; do
  timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 6 --env-groups 1 --tune $t > "gpurun_out/${T}_ab_$t.json" 2>/dev/null
  python - "$t" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_45_ab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p','grid','sort')})
PY
done
