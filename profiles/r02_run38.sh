#!/bin/bash
# round 2, call 40 (8 GPUs): default bench at N=8 with both forms of the fused APG update timed beside NCCL
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_38
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench8.out 2> gpurun_out/${T}_bench8.err
grep '"metric"' gpurun_out/${T}_bench8.out > gpurun_out/${T}_bench8.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_38_bench8.json').read().strip().splitlines()[-1])
print('N=8 ms', d['ms_per_step'], 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'e2e ms', d['e2e']['ms_per_step'], 'host GB/s', d['e2e'].get('host_gbs_all_ranks'))
print(d.get('apg_update'))
PY
