"""Size-independent properties at BASELINE.json's full sizes (configs[1]: 50 625 particles/env x 32 envs/GPU, S = 16;
configs[3]: 128 envs x 512 cloth nodes), where the CPU oracle would take minutes: sortedness, partition of unity,
env independence, the adjoint's response to a uniform perturbation, taped == recompute."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, DENSITY = 32, 3.9


def _scene(p2g_mode=0, adjoint="recompute", gravity=None, init_y=0.07):
    from unidom_b200 import confs
    from unidom_b200.mpm_simulator import SimpleMPMSimulator, create_primitive
    conf = confs.shape_elasto_plastic_conf()
    if gravity is not None:
        conf.gravity = gravity
    sim = SimpleMPMSimulator(conf, B, p2g_mode=p2g_mode, adjoint=adjoint)
    st = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.2, 0.06, 0.12], init_pos=[0.25, init_y, 0.25],
                     z_rotation_angle=0, material=2, density=DENSITY)
    st.primitives.append(create_primitive(conf, friction=0.1, softness=666, color=[0.5] * 3, size=[0.015, 0.06, 0.015],
                                          init_pos=[0.25, 0.01, 0.20]))
    return conf, sim, sim.reset_jax(st)


def test_full_size_binning_is_a_stable_sort(built_lib):
    """ud_mpm_sort_bins on 32 x 50 625 particles: perm is a permutation of each env, keys are non-decreasing along it,
    equal keys keep their input order (stable), base is int32(x * inv_dx - 0.5) without FMA contraction."""
    conf, sim, st = _scene()
    n = st.x.shape[1]
    assert n == 50625
    g = torch.Generator(device="cuda").manual_seed(0)
    x = st.x + 0.004 * torch.randn(st.x.shape, generator=g, device="cuda")           # break the lattice
    base, key, perm = sim.sort_bins(x)
    p = perm.long()
    assert torch.equal(torch.sort(p, dim=1).values, torch.arange(n, device="cuda").expand(B, n))
    ks = torch.gather(key.long(), 1, p) if key.shape == perm.shape else None
    # `key` is returned in sorted order by the ABI (keys of the sorted slots): accept either convention
    sorted_keys = key.long() if bool((key[:, 1:] >= key[:, :-1]).all()) else ks
    assert bool((sorted_keys[:, 1:] >= sorted_keys[:, :-1]).all())
    same = sorted_keys[:, 1:] == sorted_keys[:, :-1]
    assert bool((p[:, 1:][same] > p[:, :-1][same]).all())                             # stable
    ref = (x * np.float32(conf.inv_dx) - np.float32(0.5)).to(torch.int32)           # torch: mul then sub, no FMA
    assert torch.equal(base, ref)


def test_full_size_uniform_translation(built_lib):
    """Partition of unity at full size: without gravity and away from the colliders, a block moving with one velocity
    keeps it (G2P o P2G returns a uniform field, C' = 0, F stays I) and advances by S dt v; and the adjoint answers a
    UNIFORM velocity perturbation with sum_p dL/dv_p = S dt sum_p c_p for L = sum_p c_p . x'_p."""
    conf, sim, st = _scene(gravity=(0.0, 0.0, 0.0), init_y=0.16)
    n = st.x.shape[1]
    v0 = torch.tensor([0.03, -0.02, 0.05], device="cuda")
    v = v0.expand(B, n, 3).contiguous().requires_grad_(True)
    x = st.x.detach().clone().requires_grad_(True)
    action = torch.zeros((B, 6), device="cuda")
    out, _ = sim.step_jax(st._replace(x=x, v=v), action)
    S, dt = conf.steps, conf.dt
    assert float((out.v - v0).abs().max()) < 5e-6                     # 16 x (p / m) roundings on |v| = 0.06
    assert float((out.x - (st.x + S * dt * v0)).abs().max()) < 1e-6   # a few ulps of x ~ 0.3
    assert float(out.C.abs().max()) < 1e-3 and float((out.F - st.F).abs().max()) < 2e-5   # |v| / dx = 5.8: C at 1e-4 relative;
    # F = U clip(s) Vt of 16 re-factorisations of (I + dt C) F
    c = torch.randn((B, n, 3), generator=torch.Generator().manual_seed(1)).cuda() * 1e-3
    gx, gv = torch.autograd.grad((out.x * c).sum(), [x, v])
    want = S * dt * c.sum(1)
    assert float((gv.sum(1) - want).abs().max()) < 2e-3 * float(want.abs().max())
    assert float((gx.sum(1) - c.sum(1)).abs().max()) < 2e-3 * float(c.sum(1).abs().max())   # uniform shift of x


def test_full_size_envs_are_independent_and_taped_equals_recompute(built_lib):
    """Deterministic P2G at full size: identical envs give bit-identical states (no cross-env traffic), perturbing one
    env changes no other, and the taped forward is bit-identical to the plain one."""
    from unidom_b200 import _lib
    conf, sim, st = _scene(p2g_mode=_lib.UD_P2G_DETERMINISTIC)
    action = torch.tensor([0.5, 0.0, 0.6, 0.0, 0.0, 0.0], device="cuda").repeat(B, 1)
    with torch.no_grad():
        out, _ = sim.step_jax(st, action)
        for k in ("x", "v", "C", "F"):
            t = getattr(out, k)
            assert torch.equal(t[1:], t[:1].expand_as(t[1:])), k
        x2 = st.x.clone()
        x2[5] += 1e-3
        a2 = action.clone()
        a2[5, 0] = -0.3
        out2, _ = sim.step_jax(st._replace(x=x2), a2)
        keep = [b for b in range(B) if b != 5]
        for k in ("x", "v", "C", "F"):
            assert torch.equal(getattr(out2, k)[keep], getattr(out, k)[keep]), k
        assert not torch.equal(out2.x[5] - 1e-3, out.x[5])
    sim.adjoint = "tape"
    xr = st.x.detach().clone().requires_grad_(True)
    out_t, _ = sim.step_jax(st._replace(x=xr), action)
    assert sim.last_adjoint == "tape"
    for k in ("x", "v", "C", "F", "J"):
        assert torch.equal(getattr(out_t, k), getattr(out, k)), k
    (gx,) = torch.autograd.grad((out_t.x * out_t.v).sum(), [xr])
    assert torch.isfinite(gx).all() and float(gx.abs().max()) > 0


def test_full_size_cloth_envs_are_independent(built_lib):
    """configs[3] size (128 envs x 512 nodes, 50 substeps): identical envs stay bit-identical, a different action in one
    env leaves the others untouched, in the forward and in the adjoint."""
    from unidom_b200 import confs
    from unidom_b200.cloth_simulator import ClothSimulator
    conf = confs.ClothConf()
    Bc = 128
    sim = ClothSimulator(conf, Bc, None, confs.fold_cloth_mask(conf))
    st = sim.reset_jax()
    st = st._replace(primitive0=torch.cat([st.x[:, 40], torch.full((Bc, 1), 0.01, device="cuda")], dim=1))
    act = torch.tensor([0.02, 0.5, 0.1, 0.0, 0.0, 0.0, 0.0, 1.0], device="cuda").repeat(Bc, 1)   # gripper 0 drags its nodes

    def run(a):
        x = st.x.detach().clone().requires_grad_(True)
        ar = a.detach().clone().requires_grad_(True)
        o, _ = sim.step_jax(st._replace(x=x), ar)
        gx, ga = torch.autograd.grad((o.x[..., 1]).sum(), [x, ar])
        return o, gx, ga
    o1, gx1, ga1 = run(act)
    assert torch.equal(o1.x[1:], o1.x[:1].expand_as(o1.x[1:])) and torch.equal(gx1[1:], gx1[:1].expand_as(gx1[1:]))
    a2 = act.clone()
    a2[77, 1] = -0.4
    o2, gx2, ga2 = run(a2)
    keep = [b for b in range(Bc) if b != 77]
    assert torch.equal(o2.x[keep], o1.x[keep]) and torch.equal(gx2[keep], gx1[keep]) and torch.equal(ga2[keep], ga1[keep])
    assert not torch.equal(o2.x[77], o1.x[77])
