"""shape_rope at the reference's size (582 plastic particles, 30 sub-actions x 133 substeps per env step,
core/envs/shape_rope_env.py:95-123): the launch-bound case.  Times one differentiated env-step scan (forward + adjoint of
30 step_jax calls) eagerly and as two CUDA graphs (graphs.GraphedMPMScanGrad)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch

from unidom_b200 import confs
from unidom_b200.graphs import GraphedMPMScanGrad
from unidom_b200.mpm_simulator import SimpleMPMSimulator

B, T = int(os.environ.get("B", 4)), 30
conf = confs.shape_rope_conf()
sim = SimpleMPMSimulator(conf, B, device="cuda:0", env_groups=1)
st = confs.build_shape_rope(sim, density=3)
n = st.x.shape[1]
acts = torch.zeros((T, B, 6), device="cuda")
acts[:, :, 2] = 0.01 / 30 * 50          # a slow push along z
w = torch.linspace(-1, 1, st.x.numel(), device="cuda").reshape(st.x.shape) * 1e-3


def run(fn, reps):
    best = None
    for _ in range(reps):
        x = st.x.clone().requires_grad_(True)
        a = acts.clone().requires_grad_(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(x, a)
        gx, ga = torch.autograd.grad((out.x * w).sum(), [x, a])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out, gx


def eager(x, a):
    s = st._replace(x=x)
    for t in range(T):
        s, _ = sim.step_jax(s, a[t])
    return s


t_eager, o_e, g_e = run(eager, 3)
t0 = time.perf_counter()
scan = GraphedMPMScanGrad(sim, st, acts)
t_capture = time.perf_counter() - t0
t_graph, o_g, g_g = run(lambda x, a: scan(st._replace(x=x), a), 5)
sub = T * conf.steps
res = {"scene": f"shape_rope, {n} particles x {B} envs, {T} sub-actions x {conf.steps} substeps = {sub} substeps per env step",
       "eager_ms_per_env_step_fwd_bwd": t_eager * 1e3, "graph_ms_per_env_step_fwd_bwd": t_graph * 1e3,
       "eager_us_per_substep": t_eager * 1e6 / sub, "graph_us_per_substep": t_graph * 1e6 / sub,
       "capture_s": t_capture, "max_abs_dx_graph_vs_eager": float((o_g.x - o_e.x).abs().max()),
       "grad_rel_graph_vs_eager": float((g_g - g_e).abs().max() / (g_e.abs().max() + 1e-30))}
print(json.dumps(res))
