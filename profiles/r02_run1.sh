#!/bin/bash
# round 2, call 1: parity suite on the new tile layout + warp-local kernels, then A/B of the staging variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_fullsize_parity_gpu.py 2>&1 | tail -40 > gpurun_out/r02_1_tests.log
timeout 900 python -m pytest tests/test_fullsize_parity_gpu.py tests/test_clothenv.py -m gpu -q -s 2>&1 | grep -v Warning | tail -120 > gpurun_out/r02_1_fullsize.log
for t in warp=1 warp=3 warp=0; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --tune $t > gpurun_out/r02_1_bench_$t.json 2> gpurun_out/r02_1_bench_$t.err
done
tail -5 gpurun_out/r02_1_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_1_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
