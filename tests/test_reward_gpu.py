"""GPU parity of the reward kernels (csrc/reward.cu: calc_chamfer util.py:138-153, calc_l2 :156-159, with adjoints)
against the CPU oracle restatement and the reference's own calc_chamfer output in tests/golden/ref_clothenv_*.npz."""
import numpy as np
import pytest
import torch

import util
from oracle import reward as orw

pytestmark = pytest.mark.gpu


def _grad(fn, x, y, w):
    x = x.detach().clone().requires_grad_(True)
    out = fn(x, y)
    (g,) = torch.autograd.grad((out * w.to(out.device)).sum(), [x])
    return out.detach(), g


@pytest.mark.parametrize("B,P,Q", [(3, 700, 1300), (2, 1025, 1), (1, 1, 2049), (5, 512, 512)])
def test_chamfer_matches_oracle(built_lib, B, P, Q):
    """Random clouds; P, Q straddle the 256-thread CTA and the 1024-point shared-memory tile."""
    from unidom_b200 import envs
    g = torch.Generator().manual_seed(B * 1000 + P)
    x = torch.rand((B, P, 3), generator=g)
    y = torch.rand((Q, 3), generator=g)
    w = torch.randn(B, generator=g)
    ref, gref = _grad(orw.calc_chamfer, x.double(), y.double(), w.double())
    got, ggot = _grad(envs.calc_chamfer, x.cuda(), y.cuda(), w)
    assert util.rel_err(got, ref) < 2e-6
    assert util.rel_err(ggot, gref) < 2e-5 and util.cosine(ggot, gref) > 0.999999


def test_chamfer_ties_share_the_cotangent_like_jnp_min(built_lib):
    """Dyadic lattices make exactly equidistant neighbours: jnp.min's VJP (and torch.amin) split the cotangent evenly
    over them, torch.min(dim) would not.  x = 8x8 lattice, y = the same lattice shifted by half a cell in x."""
    from unidom_b200 import envs
    i, j = torch.meshgrid(torch.arange(8.0), torch.arange(8.0), indexing="ij")
    lat = torch.stack([i.flatten() / 8, torch.zeros(64), j.flatten() / 8], dim=1)
    x = lat[None].repeat(2, 1, 1)
    x[1] += torch.tensor([1 / 64, 0.0, 1 / 32])                       # env 1: unique neighbours
    y = lat + torch.tensor([1 / 16, 0.0, 0.0])
    w = torch.tensor([1.0, -2.0])
    ref, gref = _grad(orw.calc_chamfer, x, y, w)
    got, ggot = _grad(envs.calc_chamfer, x.cuda(), y.cuda(), w)
    single = torch.sqrt(((x[:, :, None] - y[None, None]) ** 2).mean(-1))
    xs = x.clone().requires_grad_(True)
    ds = torch.sqrt(((xs[:, :, None] - y[None, None]) ** 2).mean(-1))
    (g_single,) = torch.autograd.grad(((ds.min(-1).values.mean(1) + ds.min(-2).values.mean(1)) * w).sum(), [xs])
    assert int((single[0] == single[0].min(-1, keepdim=True).values).sum(-1).max()) == 2    # ties exist in env 0
    assert util.rel_err(g_single, gref) > 1e-2                        # ...and the tie rule matters
    assert util.rel_err(got, ref) < 1e-6
    assert util.rel_err(ggot, gref) < 1e-5, util.rel_err(ggot, gref)


@pytest.mark.parametrize("name", ["ref_clothenv_ep1", "ref_clothenv_ep3"])
def test_chamfer_matches_reference_output(built_lib, name):
    """`chamfer0` in the fixture is the reference's own calc_chamfer(state.x, env.goal) (oracle/gen_golden.py)."""
    from unidom_b200 import envs
    import os
    d = np.load(os.path.join(util.GOLD, name + ".npz"))
    x, goal = torch.as_tensor(d["in_x"]), torch.as_tensor(d["goal"])
    got = envs.calc_chamfer(x.cuda(), goal.cuda())
    assert util.rel_err(got, torch.as_tensor(d["chamfer0"])) < 1e-6


def test_l2_matches_oracle(built_lib):
    from unidom_b200 import envs
    g = torch.Generator().manual_seed(4)
    B, P = 3, 1000
    x, y = torch.rand((B, P, 3), generator=g), torch.rand((P, 3), generator=g)
    w = torch.randn(B, generator=g)
    ref, gref = _grad(orw.calc_l2, x.double(), y.double(), w.double())
    got, ggot = _grad(envs.calc_l2, x.cuda(), y.cuda(), w)
    assert util.rel_err(got, ref) < 1e-6 and util.rel_err(ggot, gref) < 1e-5
    got1 = envs.calc_l2(x.cuda(), y[:1].cuda())                         # (1,3) goal broadcasts (MPMEnv default)
    assert util.rel_err(got1, orw.calc_l2(x, y[:1])) < 1e-6


def test_chamfer_full_size_translation_property(built_lib):
    """BASELINE configs[1] size (50 625 points per env): a cloud against a copy of itself shifted by less than half
    the lattice spacing has chamfer 2 * sqrt(mean(shift^2)) and gradient sign(shift)-directed, env by env."""
    from unidom_b200 import envs
    n = (75, 15, 45)
    ax = [torch.arange(k, dtype=torch.float32) / 96 for k in n]
    a, b, c = torch.meshgrid(*ax, indexing="ij")
    y = torch.stack([a, b, c], -1).reshape(-1, 3).cuda()
    shift = torch.tensor([[1e-3, 0.0, 0.0], [0.0, -2e-3, 1e-3], [1e-3, 1e-3, 1e-3], [0.0, 0.0, 0.0]]).cuda()
    x = (y[None] + shift[:, None]).requires_grad_(True)
    ch = envs.calc_chamfer(x, y)
    want = 2 * torch.sqrt((shift ** 2).mean(-1))
    assert util.rel_err(ch, want) < 1e-4, (ch, want)
    (gx,) = torch.autograd.grad(ch[:3].sum(), [x])
    P = y.shape[0]
    gwant = 2 * shift[:3] / (3 * torch.sqrt((shift[:3] ** 2).mean(-1, keepdim=True))) / P
    assert util.rel_err(gx[:3].sum(1), gwant * P) < 1e-3
    assert torch.isnan(gx[3]).all() or float(gx[3].abs().max()) == 0.0   # coincident clouds: 0 * inf, as in the reference


def test_apg_update_kernels_match_the_host_logic(built_lib):
    """ud_apg_scrub_clip + ud_adam_step (apg.py:233-240, 260-267 on the flat buffer) against the torch host logic that
    the gloo CPU tests pin: NaN/inf scrub, clip only when the norm exceeds max_grad_norm, three Adam steps."""
    from unidom_b200 import apg
    g = torch.Generator().manual_seed(3)
    n = 925964                                            # the fold_cloth policy
    for scale, max_norm in ((1e-4, 0.3), (5e-3, 0.3)):     # below / above the clip
        grad = torch.randn(n, generator=g) * scale
        grad[17] = float("nan")
        grad[99] = float("inf") if scale < 1e-3 else 0.0   # inf -> FLT_MAX -> the norm overflows -> everything scaled to ~0
        ref, ref_norm = apg.reduce_policy_gradient(grad.clone(), max_norm)
        got, got_norm = apg.reduce_policy_gradient(grad.clone().cuda(), max_norm)
        assert torch.isfinite(got).all()
        if torch.isfinite(ref_norm):
            assert abs(float(got_norm) - float(ref_norm)) < 1e-5 * float(ref_norm)
            assert util.rel_err(got, ref) < 1e-6
        else:
            assert not torch.isfinite(got_norm) and float(got.abs().max()) == float(ref.abs().max()) == 0.0
    params = torch.randn(n, generator=g)
    o_cpu, o_gpu = apg.Adam(n, 1e-4, "cpu"), apg.Adam(n, 1e-4, "cuda")
    p_cpu, p_gpu = params.clone(), params.clone().cuda()
    for it in range(3):
        grad = torch.randn(n, generator=g) * 1e-3
        p_cpu = o_cpu.step(p_cpu, grad)
        p_gpu = o_gpu.step(p_gpu, grad.cuda())
    assert util.rel_err(p_gpu, p_cpu) < 1e-7 and float((p_gpu.cpu() - params).abs().max()) > 1e-5
    assert util.rel_err(o_gpu.m, o_cpu.m) < 1e-6 and util.rel_err(o_gpu.v, o_cpu.v) < 1e-6
