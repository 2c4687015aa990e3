"""Generates tests/golden/cloth_contact_state.pt: the oracle's cloth state 44 substeps into a violent
sub-action (ground contacts, velocity clipping, a closed gripper), used as the starting point of
teacher-forced adjoint windows.  Uses only oracle/ (run from the repo root)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cloth as oc  # noqa: E402

conf = oc.ClothConf()
mask = oc.fold_cloth_mask(conf)
sim = oc.ClothSim(conf, mask)
B, seed = 2, 2
st = sim.reset(B)
g = torch.Generator().manual_seed(seed)
x = st.x + 0.002 * torch.randn(st.x.shape, generator=g)
x[..., 1] = (x[..., 1].abs() * 3 + 0.01 * torch.rand(x[..., 1].shape, generator=g)) * (torch.rand(x[..., 1].shape, generator=g) > 0.5)
v = 0.05 * torch.randn(st.v.shape, generator=g)
p0 = torch.cat([x[:, 100], torch.full((B, 1), 0.02)], dim=1)
p1 = torch.cat([x[:, 300] + 0.004, torch.full((B, 1), 0.015)], dim=1)
stiff = 900.0 + 300 * torch.rand(B, generator=g)
mu = 0.3 + 0.4 * torch.rand(B, generator=g)
st = st._replace(x=x, v=v, primitive0=p0, primitive1=p1, stiffness=stiff, mu=mu)
a = torch.tensor([0.3, 0.5, -0.2, 0.0, -0.1, 0.2, 0.4, 0.3])
s = oc.index_state(st, 0)
s = s._replace(action0=torch.cat([a[:3].clamp(-2, 2) / 50, a[3:4]]), action1=torch.cat([a[4:7].clamp(-2, 2) / 50, a[7:8]]))
with torch.no_grad():
    for k in range(44):
        s = sim.step(s)
print("ground nodes:", int((s.x[:, 1] <= 1e-8).sum()), "max |v|:", float(s.v.abs().max()))
torch.save({k: getattr(s, k) for k in s._fields}, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cloth_contact_state.pt"))
