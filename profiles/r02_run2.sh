#!/bin/bash
# round 2, call 2: full parity suite with the fused G2P+P2G boundary kernel, fuse on/off A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_fullsize_parity_gpu.py 2>&1 | grep -v Warning | tail -60 > gpurun_out/r02_2_tests.log
timeout 900 python -m pytest tests/test_fullsize_parity_gpu.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert" > gpurun_out/r02_2_fullsize.log
timeout 600 python -m pytest tests/test_clothenv.py -m gpu -q -s 2>&1 | grep -E "^clothenv|^para|passed|failed|Error|assert" > gpurun_out/r02_2_clothenv.log
for t in fuse=1 fuse=0; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --tune $t > gpurun_out/r02_2_bench_$t.json 2> gpurun_out/r02_2_bench_$t.err
done
tail -8 gpurun_out/r02_2_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_2_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
