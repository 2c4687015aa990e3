#!/bin/bash
# round 2, call 4: parity suite on the new SVD / hoisted loads / hierarchical G2P, A/B against the one-sided SVD build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_4
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_fullsize_parity_gpu.py 2>&1 | grep -v Warning | tail -40 > gpurun_out/${T}_tests.log
timeout 900 python -m pytest tests/test_fullsize_parity_gpu.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert" > gpurun_out/${T}_fullsize.log
timeout 600 python -m pytest tests/test_clothenv.py -m gpu -q -s 2>&1 | grep -E "^clothenv|^para|passed|failed|Error|assert" > gpurun_out/${T}_clothenv.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_new.json 2> gpurun_out/${T}_bench_new.err
UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200_hestenes.so timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_hestenes.json 2> gpurun_out/${T}_bench_hestenes.err
tail -8 gpurun_out/${T}_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_4_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
