#!/bin/bash
# round 2, call 25: flush as one walk over the 32 staging rows with a segment-end mask (vs a row loop per segment);
# resident CTAs per SM of the persistent kernels 4 vs 3 with 2 / 4 env groups (co-residency with the other group's kernels)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_24
run() {  # name lib args
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_24_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p','grid','grid_bwd')})
PY
}
for i in 1 2; do
  run base$i _base "--env-groups 1"
  run rows$i "" "--env-groups 1"
done
run base_g2 _base ""
run rows_g2 "" ""
run rows_g2_c3 "" "--tune pers_ctas=3"
run rows_g4_c3 "" "--env-groups 4 --tune pers_ctas=3"
run rows_g2_c2 "" "--tune pers_ctas=2"
timeout 900 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_fullsize_gpu.py tests/test_fullsize_parity_gpu.py tests/test_mpmenv_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
