#!/bin/bash
# round 2, call 39 (2 GPUs): reduce-scatter + broadcast form of the fused APG update (forced at N=2) against the all-read
# form, bit for bit; both forms timed by the bench leg
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_37
timeout 400 python -m pytest tests/test_apg_gpu.py -m gpu -q -s -k "two_ranks or world1" 2>&1 | grep -vi warning | tail -30 > gpurun_out/${T}_tests.log
tail -16 gpurun_out/${T}_tests.log | cut -c1-220
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench2.out 2> gpurun_out/${T}_bench2.err
grep '"metric"' gpurun_out/${T}_bench2.out | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('apg_update'))"
