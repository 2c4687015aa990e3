#!/bin/bash
# round 2, call 29: recompute pass with two alternating raw grids + the raw {p, m} of the listed cells kept in list order
# (no 600 MB memset per call), last G2P of the recompute pass skipped; full MPM test set with the measured floors
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_28
run() {  # name lib args
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_28_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), d['peak_hbm_bytes'], {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p','grid','grid_bwd','fk','sort','gather','memset','unsort','finish_bwd')})
PY
}
for i in 1 2; do
  run base$i _base "--env-groups 1"
  run new$i "" "--env-groups 1"
done
run base_g2 _base ""
run new_g2 "" ""
run new_det "" "--p2g-mode 1"
timeout 1200 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_fullsize_gpu.py tests/test_fullsize_parity_gpu.py tests/test_mpmenv_gpu.py tests/test_multidevice_gpu.py -q -s -m gpu -k "not shape_rope_env" > gpurun_out/${T}_tests.log 2>&1; tail -5 gpurun_out/${T}_tests.log
