"""Debug: forward / backward of the mini scene with mixed materials, flag on/off (each case in its own process)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import util
from unidom_b200 import confs, _lib
from unidom_b200.mpm_simulator import SimpleMPMSimulator

case, flag, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
conf = confs.shape_elasto_plastic_conf(); conf.steps = steps; conf.n_primitive = 1
sim = SimpleMPMSimulator(conf, 2)
st = util.mini_plasticine(sim, 2, seed=9, material=2)
n = sim.material.shape[0]
rs = np.random.RandomState(3)
mats = {"mix012": rs.randint(0, 3, n), "mix02": rs.randint(0, 2, n) * 2, "mix01": rs.randint(0, 2, n), "all0": np.zeros(n), "halves": (np.arange(n) < n // 2) * 2}[case]
sim.material = torch.from_numpy(mats.astype(np.int32))
sim._material_dev = sim.material.cuda().contiguous()
if not flag:
    sim._liq_cache = (id(sim.material), False)
print(case, "flag", flag, "p2g_mode", hex(sim.params().p2g_mode), "n", n, flush=True)
act = torch.zeros((2, 6), device="cuda")
x = st.x.clone().requires_grad_(True)
out, _ = sim.step_jax(st._replace(x=x), act)
torch.cuda.synchronize()
print("  forward ok", float(out.x.abs().max()), float(out.F.abs().max()), flush=True)
(g,) = torch.autograd.grad((out.x * out.v).sum(), [x])
torch.cuda.synchronize()
print("  backward ok", float(g.abs().max()), flush=True)
