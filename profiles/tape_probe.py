"""Per-step device + wall time of the taped adjoint (development probe)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from unidom_b200 import confs
from unidom_b200.mpm_simulator import SimpleMPMSimulator
conf = confs.shape_elasto_plastic_conf()
sim = SimpleMPMSimulator(conf, 32, adjoint=sys.argv[1] if len(sys.argv) > 1 else "tape")
state = bench.build_scene(sim, bench.DENSITY)
action = torch.tensor([0.003, 0.0, 0.004, 0.0, 0.0, 0.0]).repeat(32, 1).cuda()
cot = bench.make_cotangents(state, 1)
with torch.no_grad():
    for _ in range(8):
        state, _ = sim.step_jax(state, action)
state = bench.detach_state(state)
rows = []
for i in range(24):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out, grads, _ = bench.fwd_bwd(sim, state, action, cot)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    rows.append((e0.elapsed_time(e1), (t1 - t0) * 1e3, torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
for r in rows:
    print("dev %.2f ms  cpu-enqueue %.2f ms  alloc %.2f GB reserved %.2f GB" % r)
