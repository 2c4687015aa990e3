#!/bin/bash
# round 2, call 36 (2 GPUs): bench at N=2 with the APG leg behind its watchdog (line printed even if the leg cannot finish)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r02_34
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench2.out 2> gpurun_out/${T}_bench2.err
echo "rc=$?"
grep '"metric"' gpurun_out/${T}_bench2.out > gpurun_out/${T}_bench2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_34_bench2.json').read().strip().splitlines()[-1])
print('N=2 ms', d['ms_per_step'], 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'e2e ms', d['e2e']['ms_per_step'], 'host GB/s', d['e2e'].get('host_gbs_all_ranks'), d['e2e'].get('numa'))
print(d.get('apg_update'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${T}_ref2.out 2> gpurun_out/${T}_ref2.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/${T}_ref2.out
tail -3 gpurun_out/${T}_bench2.err
