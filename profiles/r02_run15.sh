#!/bin/bash
# round 2, call 15: envs of a call as 1 / 2 / 4 sub-batches on separate streams, persistent vs one-tile P2G under concurrency
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_15
timeout 600 python -m pytest tests/test_jaxrng_cpu.py tests/test_mpmenv_gpu.py tests/test_mpm_gpu.py -q 2>&1 | tail -3
for g in 1 2 4; do for t in pers=1 pers=0; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --env-groups $g --tune $t > gpurun_out/${T}_bench_g${g}_$t.json 2> gpurun_out/${T}_bench_g${g}_$t.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_15_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), 'fwd-only %.3e'%d['forward_only']['value'], 'taped', round(d['taped']['ms_per_step'],3))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/${T}_bench_g2_pers=1.err
