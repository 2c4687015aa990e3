#!/bin/bash
# round 2, call 10: full parity suite + bench after: persistent P2G / P2G^T with next-tile prefetch, forward grid update skipping inactive primitives
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_10
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_fullsize_parity_gpu.py 2>&1 | grep -v Warning | tail -30 > gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 900 python -m pytest tests/test_fullsize_parity_gpu.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert" > gpurun_out/${T}_fullsize.log
tail -2 gpurun_out/${T}_fullsize.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_10_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
