"""GPU parity of the mass-spring cloth sub-action (50 substeps, fwd + adjoint) against the CPU oracle."""
import numpy as np
import pytest
import torch

import util
from oracle import cloth as oc

pytestmark = pytest.mark.gpu


def _scene(B, seed, float_stiffness, substeps=50):
    from unidom_b200.cloth_simulator import ClothSimulator
    conf = oc.ClothConf()
    mask = oc.fold_cloth_mask(conf)
    sim = ClothSimulator(conf, B, None, mask)
    sim.SUBSTEPS = substeps
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


@pytest.mark.parametrize("float_stiffness,substeps", [(False, 50), (True, 50), (True, 5)])
def test_cloth_forward_parity(built_lib, float_stiffness, substeps):
    """50 substeps = one reference sub-action (chaotic: compared against the oracle's own fp32-vs-fp64
    floor, SURVEY hard part 2); 5 substeps = teacher-forced window with a tight tolerance."""
    B = 3
    conf, mask, sim, st, act = _scene(B, 1, float_stiffness, substeps)
    out, _ = sim.step_jax(st, act)
    osim = oc.ClothSim(conf, mask)
    with torch.no_grad():
        ref = oc.step_batch(osim, _to_oracle(st), act.cpu(), substeps)
        ref64 = oc.step_batch(oc.ClothSim(conf, mask, torch.float64), _to_oracle(st, torch.float64),
                              act.cpu().double(), substeps)
    for k in ("x", "v", "primitive0", "primitive1", "action0", "action1"):
        e = util.rel_err(getattr(out, k), getattr(ref, k))
        fl = util.rel_err(getattr(ref, k), getattr(ref64, k))
        print(f"cloth {k}: cuda-vs-oracle32 {e:.3e}  oracle32-vs-64 {fl:.3e}")
        assert e < (max(1e-4, 5 * fl) if substeps == 50 else max(2e-5, 3 * fl)), (k, e, fl)


@pytest.mark.parametrize("float_stiffness,substeps", [(False, 50), (True, 50), (True, 3)])
def test_cloth_backward_parity(built_lib, float_stiffness, substeps):
    """The cloth cotangent is re-normalised 8x per substep (norm_grad) on chaotic dynamics, so over 50
    substeps two correct fp32 implementations only agree to the oracle's fp32-vs-fp64 floor (printed);
    the 3-substep window is the tight check of the adjoint formulas."""
    B = 2
    conf, mask, sim, st, act = _scene(B, 2, float_stiffness, substeps)
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
    ref = run(lambda s, a: oc.step_batch(osim, s, a, substeps), _to_oracle(st), act.cpu(), lambda t: t)
    osim64 = oc.ClothSim(conf, mask, torch.float64)
    ref64 = run(lambda s, a: oc.step_batch(osim64, s, a, substeps), _to_oracle(st, torch.float64),
                act.cpu().double(), lambda t: t)
    pert = None
    if substeps == 50:
        # sensitivity floor: the oracle's own gradient after perturbing the input positions by 1e-6
        # (relative) -- a few fp32 ulps decide on which substep a node touches the ground / leaves a clip
        gp = torch.Generator().manual_seed(11)
        ost = _to_oracle(st)
        ost = ost._replace(x=ost.x * (1 + 1e-6 * torch.randn(ost.x.shape, generator=gp)))
        pert = run(lambda s, a: oc.step_batch(osim, s, a, substeps), ost, act.cpu(), lambda t: t)
    for k in ref:
        e = util.rel_err(got[k], ref[k])
        fl = util.rel_err(ref[k], ref64[k])
        big = float(ref[k].abs().max()) > 1e-20
        cs = util.cosine(got[k], ref[k]) if big else 1.0
        csf = util.cosine(ref[k], ref64[k]) if big else 1.0
        line = (f"cloth[{substeps}] grad {k:12s} rel {e:.3e} cos {cs:.6f} | oracle32-vs-64 rel {fl:.3e} cos {csf:.6f} "
                f"max|ref| {float(ref[k].abs().max()):.3e}")
        if pert is not None and big:
            fp_, csp = util.rel_err(pert[k], ref[k]), util.cosine(pert[k], ref[k])
            line += f" | oracle32 under 1e-6 input perturbation rel {fp_:.3e} cos {csp:.6f}"
            fl, csf = max(fl, fp_), min(csf, csp)
        print("\n" + line)
        if big:
            if substeps == 50:
                assert cs >= min(0.999, 1 - 3 * (1 - csf)), (k, cs, csf)
                assert e < max(1e-3, 3 * fl), (k, e, fl)
            else:
                assert cs >= 0.99999 and e < 1e-3, (k, cs, e)


@pytest.mark.parametrize("substeps", [1, 4])
def test_cloth_adjoint_window_from_contact_state(built_lib, substeps):
    """Teacher-forced adjoint windows starting from the oracle's state 44 substeps into a violent
    sub-action (tests/golden/cloth_contact_state.pt: ground contacts, |v| clipped at max_v, closed
    gripper): every discrete branch of the step is active, and the window is short enough that both
    implementations take the same branches."""
    import os
    from unidom_b200.cloth_simulator import ClothSimulator, ClothState
    d = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cloth_contact_state.pt"))
    conf = oc.ClothConf()
    mask = oc.fold_cloth_mask(conf)
    st_o = oc.ClothState(*[d[k][None] for k in oc.ClothState._fields])
    assert int((st_o.x[0, :, 1] <= 1e-8).sum()) >= 1 and float(st_o.v.abs().max()) >= 1.99
    act = torch.cat([d["action0"][:3] * 50, d["action0"][3:], d["action1"][:3] * 50, d["action1"][3:]])[None]
    sim = ClothSimulator(conf, 1, None, mask)
    sim.SUBSTEPS = substeps
    st = ClothState(*[t.cuda() for t in st_o])
    g = torch.Generator().manual_seed(3)
    cx, cv = torch.randn(st.x.shape, generator=g), torch.randn(st.x.shape, generator=g)

    def run(step, state, action, dev):
        x = state.x.detach().clone().requires_grad_(True)
        v = state.v.detach().clone().requires_grad_(True)
        mu = state.mu.detach().clone().requires_grad_(True)
        a = action.detach().clone().requires_grad_(True)
        out = step(state._replace(x=x, v=v, mu=mu), a)
        L = (out.x * cx.to(dev)).sum() + (out.v * cv.to(dev)).sum()
        return torch.autograd.grad(L, [x, v, mu, a]), out

    got, o1 = run(lambda s, a: sim.step_jax(s, a)[0], st, act.cuda(), "cuda")
    osim = oc.ClothSim(conf, mask)
    ref, o2 = run(lambda s, a: oc.step_batch(osim, s, a, substeps), st_o, act, "cpu")
    assert util.rel_err(o1.x, o2.x) < 1e-6 and util.rel_err(o1.v, o2.v) < 1e-4
    for name, a, b in zip(("x", "v", "mu", "action"), got, ref):
        cs = util.cosine(a, b)
        e = util.rel_err(a, b)
        print(f"cloth window[{substeps}] grad {name}: rel {e:.3e} cos {cs:.12f}")
        assert cs > 0.9999999 and e < 1e-4, (name, cs, e)


def test_fused_scan_is_bit_identical_to_step_by_step(built_lib):
    """ud_cloth_multi_step_{fwd,bwd} (the scan over sub-actions of cloth_env.py:211 in one call) against the loop of
    ud_cloth_step_{fwd,bwd}: same kernels, same order => bit-identical states and gradients."""
    B, T = 3, 6
    conf, mask, sim, st, act = _scene(B, 4, True)
    dev = st.x.device
    g = torch.Generator().manual_seed(8)
    actions = (torch.rand((T, B, 8), generator=g) * 1.2 - 0.4).to(dev)
    actions[..., 3] = (actions[..., 3] > 0.3).float()
    cx = torch.randn(st.x.shape, generator=g).to(dev)

    def run(fused):
        x = st.x.detach().clone().requires_grad_(True)
        mu = st.mu.detach().clone().requires_grad_(True)
        k = st.stiffness.detach().clone().requires_grad_(True)
        a = actions.detach().clone().requires_grad_(True)
        s = st._replace(x=x, mu=mu, stiffness=k)
        if fused:
            s = sim.scan_step_jax(s, a)
        else:
            for t in range(T):
                s, _ = sim.step_jax(s, a[t])
        L = (s.x * cx).sum() + (s.v * cx).sum() + s.primitive0.sum()
        return s, torch.autograd.grad(L, [x, mu, k, a])

    s1, g1 = run(False)
    s2, g2 = run(True)
    for k in ("x", "v", "primitive0", "primitive1", "action0", "action1"):
        assert torch.equal(getattr(s1, k), getattr(s2, k)), k
    for name, a, b in zip(("x", "mu", "stiffness", "actions"), g1, g2):
        assert torch.equal(a, b), (name, float((a - b).abs().max()))
    with torch.no_grad():                      # forward-only path keeps no checkpoints
        s3 = sim.scan_step_jax(st, actions)
    assert torch.equal(s3.x, s1.x)


def _tshirt_mask(N):
    """A T-shirt-like mask on the central (N/2)^2 window (fold_cloth_tshirt_env.py:52-71 reads it from a jpg)."""
    h = N // 2
    m = np.zeros((h, h), np.int32)
    m[h // 8: h // 8 + h // 4, :] = 1                    # sleeves
    m[h // 8:, h // 4: h - h // 4] = 1                   # body
    m[h // 8: h // 8 + h // 16, h // 2 - h // 10: h // 2 + h // 10] = 0   # collar notch
    full = np.zeros((N, N), np.int32)
    full[N // 2 - h // 2: N // 2 - h // 2 + h, N // 2 - h // 2: N // 2 - h // 2 + h] = m
    return full


def _big_scene(B, N, seed, substeps):
    from unidom_b200.cloth_simulator import ClothSimulator
    conf = oc.ClothConf()
    conf.N, conf.stiffness, conf.dt, conf.mu = N, 5000, 0.5e-3, 0.9         # fold_cloth_tshirt_env.py:21-31
    mask = _tshirt_mask(N)
    sim = ClothSimulator(conf, B, None, mask)
    sim.SUBSTEPS = substeps
    st = sim.reset_jax()
    g = torch.Generator().manual_seed(seed)
    dev = st.x.device
    x = st.x.cpu() + 0.001 * torch.randn(st.x.shape, generator=g)
    x[..., 1] = (x[..., 1].abs() * 3 + 0.01 * torch.rand(x[..., 1].shape, generator=g)) * (torch.rand(x[..., 1].shape, generator=g) > 0.5)
    v = 0.05 * torch.randn(st.v.shape, generator=g)
    P = x.shape[1]
    p0 = torch.cat([x[:, P // 5], torch.full((B, 1), 0.02)], dim=1)           # gripper 0 on a node of the first CTA
    p1 = torch.cat([x[:, P - 7] + 0.002, torch.full((B, 1), 0.015)], dim=1)   # gripper 1 next to a node of the last CTA
    st = st._replace(x=x.to(dev), v=v.to(dev), primitive0=p0.to(dev), primitive1=p1.to(dev),
                     stiffness=(4000.0 + 2000 * torch.rand(B, generator=g)).to(dev), mu=(0.3 + 0.6 * torch.rand(B, generator=g)).to(dev))
    act = torch.tensor([[0.3, 0.5, -0.2, 0.0, -0.1, 0.2, 0.4, 0.3], [0.6, -0.4, 0.1, 1.0, 0.0, 0.3, 0.0, 0.0]])[:B].to(dev)
    return conf, mask, sim, st, act


def _cloth_grads(step, state, action, cot, names, todev):
    req = {k: getattr(state, k).detach().clone().requires_grad_(True) for k in names}
    a = action.detach().clone().requires_grad_(True)
    out = step(state._replace(**req), a)
    L = sum((getattr(out, k) * todev(cot[k]).to(getattr(out, k).dtype)).sum() for k in cot)
    gr = torch.autograd.grad(L, [req[k] for k in names] + [a], allow_unused=True)
    res = {k: (g_ if g_ is not None else torch.zeros_like(req[k])) for k, g_ in zip(names, gr[:-1])}
    res["action"] = gr[-1]
    return out, res


@pytest.mark.parametrize("N", [160, 180])
def test_tshirt_sized_cloth_on_a_cluster_matches_oracle(built_lib, N):
    """fold_tshirt scale (fold_cloth_tshirt_env.py: N = 180, 3 573 nodes): one thread-block
    cluster per env -- N=160 gives 3 5xx nodes on 7 CTAs x 512 threads, N=180 4 512 nodes on 5 CTAs x 1024 threads.
    Teacher-forced windows (5 substeps forward, 3 substeps adjoint) against the oracle, next to its fp32-vs-fp64 floor."""
    B = 2
    conf, mask, sim, st, act = _big_scene(B, N, 3, 5)
    P = st.x.shape[1]
    assert P > 3000 and P == int(mask.sum())
    out, _ = sim.step_jax(st, act)
    osim = oc.ClothSim(conf, mask)
    with torch.no_grad():
        ref = oc.step_batch(osim, _to_oracle(st), act.cpu(), 5)
        ref64 = oc.step_batch(oc.ClothSim(conf, mask, torch.float64), _to_oracle(st, torch.float64), act.cpu().double(), 5)
    for k in ("x", "v", "primitive0", "primitive1"):
        e = util.rel_err(getattr(out, k), getattr(ref, k))
        fl = util.rel_err(getattr(ref, k), getattr(ref64, k))
        print(f"tshirt cloth ({P} nodes) fwd {k}: rel {e:.3e} | oracle32-vs-64 {fl:.3e}")
        assert e < max(2e-5, 3 * fl), (k, e, fl)
    sim.SUBSTEPS = 3
    g = torch.Generator().manual_seed(5)
    cot = {"x": torch.randn(st.x.shape, generator=g), "v": torch.randn(st.v.shape, generator=g),
           "primitive0": torch.randn((B, 4), generator=g), "primitive1": torch.randn((B, 4), generator=g)}
    names = ["x", "v", "primitive0", "primitive1", "mu", "stiffness"]
    _, got = _cloth_grads(lambda s, a: sim.step_jax(s, a)[0], st, act, cot, names, lambda t: t.to(st.x.device))
    _, want = _cloth_grads(lambda s, a: oc.step_batch(osim, s, a, 3), _to_oracle(st), act.cpu(), cot, names, lambda t: t)
    for k in want:
        if float(want[k].abs().max()) <= 1e-20:
            continue
        e, cs = util.rel_err(got[k], want[k]), util.cosine(got[k], want[k])
        print(f"tshirt cloth adjoint {k:12s} rel {e:.3e} cos {cs:.8f}")
        assert cs >= 0.99999 and e < 1e-3, (k, cs, e)


def test_cluster_path_equals_single_cta_path(built_lib):
    """The 512-node cloth forced onto a 4-CTA cluster (ud_tuning_set cloth_cta_nodes=128): the forward has no
    reductions -> bit-identical; the adjoint differs only by the order of the per-env norm sums."""
    from unidom_b200 import _lib
    B = 3
    conf, mask, sim, st, act = _scene(B, 6, True, 50)
    g = torch.Generator().manual_seed(12)
    cot = {"x": torch.randn(st.x.shape, generator=g), "v": torch.randn(st.v.shape, generator=g),
           "primitive0": torch.randn((B, 4), generator=g), "primitive1": torch.randn((B, 4), generator=g)}
    names = ["x", "v", "primitive0", "primitive1", "mu", "stiffness"]
    dev = st.x.device
    L = _lib.lib()
    res = {}
    try:
        for nodes in (1024, 128):
            assert L.ud_tuning_set(b"cloth_cta_nodes", nodes) >= 0
            res[nodes] = _cloth_grads(lambda s, a: sim.step_jax(s, a)[0], st, act, cot, names, lambda t: t.to(dev))
    finally:
        L.ud_tuning_set(b"cloth_cta_nodes", 1024)
    (o1, g1), (o2, g2) = res[1024], res[128]
    for k in ("x", "v", "primitive0", "primitive1", "action0", "action1"):
        assert torch.equal(getattr(o1, k), getattr(o2, k)), k
    for k in g1:
        if float(g1[k].abs().max()) <= 1e-20:
            continue
        e, cs = util.rel_err(g2[k], g1[k]), util.cosine(g2[k], g1[k])
        print(f"cluster vs single CTA adjoint {k:12s} rel {e:.3e} cos {cs:.8f}")
        assert cs > 0.9999 and e < 1e-2, (k, cs, e)       # 50 chaotic substeps amplify the last-bit differences of the norms
