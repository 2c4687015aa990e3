#!/bin/bash
# round 2, call 35: mixed-material warps with UD_P2G_LIQUID_FAST (the per-lane branch deadlocked, see p2g_front_loaded)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for c in "halves 1 4" "mix02 1 1" "mix012 1 4"; do
  CUDA_LAUNCH_BLOCKING=1 timeout 45 python profiles/micro/dbg_mixed.py $c 2>&1 | grep -v "Warn\|detach\|print(" | tail -4
done > gpurun_out/r02_33_dbg.log 2>&1
cat gpurun_out/r02_33_dbg.log
grep -q "backward ok" gpurun_out/r02_33_dbg.log || exit 1
timeout 600 python -m pytest tests/test_golden_gpu.py tests/test_mpm_gpu.py tests/test_xla_ffi.py tests/test_mpmenv_gpu.py -q -s -m gpu -k "not shape_rope_env" -o faulthandler_timeout=120 > gpurun_out/r02_33_tests.log 2>&1; tail -3 gpurun_out/r02_33_tests.log
