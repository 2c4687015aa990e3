#!/bin/bash
# round 2, call 12: fused APG update (1 GPU), new bench.py paths: default config with flat e2e buffers, the other BASELINE configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_12
timeout 600 python -m pytest tests/test_apg_gpu.py tests/test_mpm_gpu.py -m gpu -q -s 2>&1 | grep -E "fused vs|passed|failed|Error|assert|skipped" | tail -20 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
for c in pour_water whip_rope cloth_para; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --config whip_rope --ckpt-window 70 > gpurun_out/${T}_bench_whip_rope_k70.json 2> gpurun_out/${T}_bench_whip_rope_k70.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_12_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, 'ms', round(d['ms_per_step'],3), 'value', '%.3e'%d['value'], 'e2e', d.get('e2e',{}).get('value') and '%.3e'%d['e2e']['value'], 'frac', r.get('step_frac_fwdbwd'), 'peak_hbm', d.get('peak_hbm_bytes'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/${T}_bench*.err
