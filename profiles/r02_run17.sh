#!/bin/bash
# round 2, call 17: segment table + 8-row flush in the warp-local scatter kernels; differentiated CUDA-graph scan
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_17
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_fullsize_parity_gpu.py 2>&1 | grep -v Warning | tail -25 > gpurun_out/${T}_tests.log
tail -6 gpurun_out/${T}_tests.log
timeout 900 python -m pytest tests/test_fullsize_parity_gpu.py -m gpu -q 2>&1 | tail -2
for t in "--env-groups 1" "--env-groups 2"; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline $t > "gpurun_out/${T}_bench_$(echo $t | tr ' =' '__').json" 2> gpurun_out/${T}_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_17_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
