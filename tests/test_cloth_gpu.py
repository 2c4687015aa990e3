"""GPU parity of the mass-spring cloth sub-action (50 substeps, fwd + adjoint) against the CPU oracle."""
import numpy as np
import pytest
import torch

import util
from oracle import cloth as oc

pytestmark = pytest.mark.gpu


def _scene(B, seed, float_stiffness):
    from unidom_b200.cloth_simulator import ClothSimulator
    conf = oc.ClothConf()
    mask = oc.fold_cloth_mask(conf)
    sim = ClothSimulator(conf, B, None, mask)
    st = sim.reset_jax()
    g = torch.Generator().manual_seed(seed)
    dev = st.x.device
    # lift part of the cloth, perturb, put gripper 0 on a node and gripper 1 next to another
    x = st.x.cpu() + 0.002 * torch.randn(st.x.shape, generator=g)
    x[..., 1] = (x[..., 1].abs() * 3 + 0.01 * torch.rand(x[..., 1].shape, generator=g)) * (torch.rand(x[..., 1].shape, generator=g) > 0.5)
    v = 0.05 * torch.randn(st.v.shape, generator=g)
    p0 = torch.cat([x[:, 100], torch.full((B, 1), 0.02)], dim=1)
    p1 = torch.cat([x[:, 300] + 0.004, torch.full((B, 1), 0.015)], dim=1)
    stiff = (900.0 + 300 * torch.rand(B, generator=g)) if float_stiffness else st.stiffness.cpu()
    mu = 0.3 + 0.4 * torch.rand(B, generator=g)
    st = st._replace(x=x.to(dev), v=v.to(dev), primitive0=p0.to(dev), primitive1=p1.to(dev),
                     stiffness=stiff.to(dev), mu=mu.to(dev))
    act = torch.tensor([[0.3, 0.5, -0.2, 0.0, -0.1, 0.2, 0.4, 0.3], [2.6, -0.4, 0.1, 1.0, 0.0, 0.0, 0.0, 0.0],
                        [0.0, 0.06, 0.0, 0.2, 0.5, 0.1, -3.0, 0.0]])[:B].to(dev)
    return conf, mask, sim, st, act


def _to_oracle(st, dtype=torch.float32):
    def cv(t):
        t = t.detach().cpu()
        return t.to(dtype) if t.is_floating_point() else t
    o = oc.ClothState(*[cv(t) for t in st])
    return o._replace(stiffness=o.stiffness.to(dtype))


@pytest.mark.parametrize("float_stiffness", [False, True])
def test_cloth_forward_parity(built_lib, float_stiffness):
    B = 3
    conf, mask, sim, st, act = _scene(B, 1, float_stiffness)
    out, _ = sim.step_jax(st, act)
    osim = oc.ClothSim(conf, mask)
    with torch.no_grad():
        ref = oc.step_batch(osim, _to_oracle(st), act.cpu())
        ref64 = oc.step_batch(oc.ClothSim(conf, mask, torch.float64), _to_oracle(st, torch.float64), act.cpu().double())
    for k in ("x", "v", "primitive0", "primitive1", "action0", "action1"):
        e = util.rel_err(getattr(out, k), getattr(ref, k))
        fl = util.rel_err(getattr(ref, k), getattr(ref64, k))
        print(f"cloth {k}: cuda-vs-oracle32 {e:.3e}  oracle32-vs-64 {fl:.3e}")
        assert e < max(1e-4, 5 * fl), (k, e, fl)


@pytest.mark.parametrize("float_stiffness", [False, True])
def test_cloth_backward_parity(built_lib, float_stiffness):
    B = 2
    conf, mask, sim, st, act = _scene(B, 2, float_stiffness)
    dev = st.x.device
    g = torch.Generator().manual_seed(3)
    cot = {"x": torch.randn(st.x.shape, generator=g), "v": torch.randn(st.v.shape, generator=g),
           "primitive0": torch.randn((B, 4), generator=g), "primitive1": torch.randn((B, 4), generator=g)}
    names = ["x", "v", "primitive0", "primitive1", "mu"] + (["stiffness"] if float_stiffness else [])

    def run(step, state, action, todev):
        req = {k: getattr(state, k).detach().clone().requires_grad_(True) for k in names}
        a = action.detach().clone().requires_grad_(True)
        out = step(state._replace(**req), a)
        L = sum((getattr(out, k) * todev(cot[k]).to(getattr(out, k).dtype)).sum() for k in cot)
        gr = torch.autograd.grad(L, [req[k] for k in names] + [a], allow_unused=True)
        res = {k: (g_ if g_ is not None else torch.zeros_like(req[k])) for k, g_ in zip(names, gr[:-1])}
        res["action"] = gr[-1]
        return res

    got = run(lambda s, a: sim.step_jax(s, a)[0], st, act, lambda t: t.to(dev))
    osim = oc.ClothSim(conf, mask)
    ref = run(lambda s, a: oc.step_batch(osim, s, a), _to_oracle(st), act.cpu(), lambda t: t)
    osim64 = oc.ClothSim(conf, mask, torch.float64)
    ref64 = run(lambda s, a: oc.step_batch(osim64, s, a), _to_oracle(st, torch.float64), act.cpu().double(), lambda t: t)
    for k in ref:
        e = util.rel_err(got[k], ref[k])
        fl = util.rel_err(ref[k], ref64[k])
        cs = util.cosine(got[k], ref[k]) if float(ref[k].abs().max()) > 0 else 1.0
        print(f"cloth grad {k:12s} rel {e:.3e} cos {cs:.6f} | oracle32-vs-64 {fl:.3e} max|ref| {float(ref[k].abs().max()):.3e}")
        if float(ref[k].abs().max()) > 1e-20:
            assert cs >= 0.999, (k, cs)
            assert e < max(1e-3, 20 * fl), (k, e, fl)
