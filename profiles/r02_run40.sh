#!/bin/bash
# round 2, call 42: pour_water after slimming the mixed-warp liquid override (k_p2g_pers<.,LIQ> had 76 bytes of spills)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_40
for c in "pour --config pour_water" "push --env-groups 1"; do
  set -- $c; name=$1; shift
  timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  python - "$name" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_40_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p','grid','grid_bwd','sort')})
PY
done
timeout 300 python -m pytest tests/test_mpm_gpu.py -q -m gpu -k "backward_parity or forward_parity" -o faulthandler_timeout=100 2>&1 | tail -2
