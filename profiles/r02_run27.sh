#!/bin/bash
# round 2, call 28: grid kernels specialised on (SDF kind, position control) with one inlined copy of the collider code
# (instruction-cache misses were their top stall), transposed warp reduction of the primitive cotangents, k_fk_bwd with
# the per-row Jacobians computed in parallel; on top of the shell-job skip and the downward G2P walk of call 27
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_27
run() {  # name lib args
  UNIDOM_B200_LIB=$PWD/unidom_b200/libunidom_b200$2.so timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline $3 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r02_27_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],3), {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items() if k in ('p2g','g2p_bwd','p2g_bwd','g2p','grid','grid_bwd','fk','sort','gather','memset','unsort','finish_bwd')})
PY
}
for i in 1 2; do
  run base$i _base "--env-groups 1"
  run new$i "" "--env-groups 1"
done
run base_g2 _base ""
run new_g2 "" ""
timeout 1200 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_fullsize_gpu.py tests/test_fullsize_parity_gpu.py tests/test_mpmenv_gpu.py -q -s -m gpu -k "not push_env and not shape_rope_env" > gpurun_out/${T}_tests.log 2>&1; tail -5 gpurun_out/${T}_tests.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --env-groups 1"
for k in k_grid_fwd k_grid_bwd; do
ncu --set full --clock-control none --import-source on -k regex:"${k}" -s 40 -c 1 -o gpurun_out/${T}_prof_$k $CMD > gpurun_out/${T}_ncu_$k.log 2>&1
done
