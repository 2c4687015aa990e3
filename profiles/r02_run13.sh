#!/bin/bash
# round 2, call 13 (2 GPUs): one process / two devices / two host threads; fused APG update over peer memory; bench at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_13
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
numactl -H > gpurun_out/${T}_numa.txt 2>&1 || lscpu | grep -i numa > gpurun_out/${T}_numa.txt
timeout 600 python -m pytest tests/test_apg_gpu.py -m gpu -q -s -k two_ranks 2>&1 | grep -vi warning | tail -40 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log | tail -25
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err
tail -1 gpurun_out/${T}_bench2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=2 ms', d['ms_per_step'], 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['e2e'].get('host_gbs_all_ranks'), d.get('apg_update'))
"
grep -ci "nvls" gpurun_out/${T}_bench2.err
tail -5 gpurun_out/${T}_bench2.err
