#!/bin/bash
# round 2, call 37 (8 GPUs): default bench at N=8 (weak scaling, e2e through host memory, APG update leg: NCCL vs fused)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_35
numactl -H > gpurun_out/${T}_numa.txt 2>&1 || lscpu > gpurun_out/${T}_numa.txt
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench8.out 2> gpurun_out/${T}_bench8.err
grep '"metric"' gpurun_out/${T}_bench8.out > gpurun_out/${T}_bench8.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_35_bench8.json').read().strip().splitlines()[-1])
print('N=8 ms', d['ms_per_step'], 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'e2e ms', d['e2e']['ms_per_step'], 'host GB/s', d['e2e'].get('host_gbs_all_ranks'), d['e2e'].get('numa'))
print(d.get('apg_update'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --config cloth_para > gpurun_out/${T}_cloth8.out 2> gpurun_out/${T}_cloth8.err
grep '"metric"' gpurun_out/${T}_cloth8.out > gpurun_out/${T}_cloth8.json; cut -c1-400 gpurun_out/${T}_cloth8.json
tail -3 gpurun_out/${T}_bench8.err gpurun_out/${T}_cloth8.err
