"""GPU parity of the MLS-MPM step (CUDA, through the C ABI) against the CPU oracle."""
import numpy as np
import pytest
import torch

import util
from oracle import mpm as omp

pytestmark = pytest.mark.gpu


def _conf(steps=8, n_primitive=1, **kw):
    from unidom_b200 import confs
    c = confs.shape_elasto_plastic_conf()
    c.steps = steps
    c.n_primitive = n_primitive
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _sim(conf, B, **kw):
    from unidom_b200.mpm_simulator import SimpleMPMSimulator
    return SimpleMPMSimulator(conf, B, use_position_control=conf.use_position_control, **kw)


def _actions(B, n_prim, seed=1):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand((B, 6 * n_prim), generator=g) * 2.4 - 1.2   # some entries beyond the [-1,1] clip
    a[:, 3:6] *= 0.3
    return a


def _oracle_step(conf, sim, state, action, dtype=torch.float32):
    osim = omp.Simulator(util.oracle_conf(conf), sim.material.clone(), sim.h.clone(), dtype=dtype)
    ost = util.to_oracle_state(state, dtype)
    return omp.step_batch(osim, ost, action.cpu().to(dtype))


def test_sort_bins_bit_exact(built_lib):
    conf = _conf()
    B = 3
    sim = _sim(conf, B)
    st = util.mini_plasticine(sim, B, seed=3)
    # throw a few particles out of the grid to exercise the clamped key
    x = st.x.clone()
    x[0, :5] = torch.tensor([-0.01, 0.4, 0.7], device=x.device)
    base, key, perm = sim.sort_bins(x)
    xs = x.cpu().numpy().astype(np.float32)
    inv_dx = np.float32(conf.inv_dx)
    ref_base = (xs * inv_dx - np.float32(0.5)).astype(np.int32)
    assert np.array_equal(base.cpu().numpy(), ref_base)
    res = np.array(conf.res)
    cb = np.clip(ref_base, 0, res - 1)
    nb = (res + 3) // 4
    blk = ((cb[..., 0] >> 2) * nb[1] + (cb[..., 1] >> 2)) * nb[2] + (cb[..., 2] >> 2)
    ref_key = (blk << 6) | ((cb[..., 0] & 3) << 4) | ((cb[..., 1] & 3) << 2) | (cb[..., 2] & 3)
    assert np.array_equal(key.cpu().numpy(), ref_key.astype(np.int32))
    ref_perm = np.stack([np.argsort(ref_key[b], kind="stable") for b in range(B)]).astype(np.int32)
    assert np.array_equal(perm.cpu().numpy(), ref_perm)


def test_sort_bins_crowded_cells_bit_exact(built_lib):
    """Cells holding thousands of particles (whip_rope at density 25: 2 025 per cell) are ranked by k_rank_big (bitmap +
    popcount scan) instead of the per-particle counting loop: same stable argsort, bit for bit.  Three layouts: particle
    indices of a cell contiguous (a lattice, like add_box), interleaved over the whole index range (spread beyond
    the bitmap -> the cooperative fallback) and a mix with ordinary sparse cells."""
    conf = _conf()
    B = 2
    rng = np.random.RandomState(11)
    n = 150000
    sim = _sim(conf, B)
    sim.n_particles = n
    dx = 1.0 / conf.inv_dx
    x = np.empty((B, n, 3), np.float32)
    # env 0: 40 crowded cells, each a contiguous index range; the rest spread thinly over the grid
    cells = rng.randint(4, 28, size=(60, 3))
    owner = np.minimum(np.arange(n) // 3000, 59)
    x[0] = (cells[owner] + 0.5 + rng.uniform(0.01, 0.98, (n, 3))) * dx
    thin = rng.rand(n) < 0.2
    x[0][thin] = rng.uniform(0.05, 0.6, (int(thin.sum()), 3))
    # env 1: the same cells but every cell's particles interleaved over the whole index range
    owner1 = rng.randint(0, 12, n)
    x[1] = (cells[owner1] + 0.5 + rng.uniform(0.01, 0.98, (n, 3))) * dx
    xt = torch.from_numpy(x).cuda()
    base, key, perm = sim.sort_bins(xt)
    inv_dx = np.float32(conf.inv_dx)
    ref_base = (x * inv_dx - np.float32(0.5)).astype(np.int32)
    assert np.array_equal(base.cpu().numpy(), ref_base)
    res = np.array(conf.res)
    cb = np.clip(ref_base, 0, res - 1)
    nb = (res + 3) // 4
    blk = ((cb[..., 0] >> 2) * nb[1] + (cb[..., 1] >> 2)) * nb[2] + (cb[..., 2] >> 2)
    ref_key = (blk << 6) | ((cb[..., 0] & 3) << 4) | ((cb[..., 1] & 3) << 2) | (cb[..., 2] & 3)
    assert np.array_equal(key.cpu().numpy(), ref_key.astype(np.int32))
    occ = np.bincount(ref_key[1])
    assert occ.max() > 5000                                                       # crowded for real
    ref_perm = np.stack([np.argsort(ref_key[b], kind="stable") for b in range(B)]).astype(np.int32)
    assert np.array_equal(perm.cpu().numpy(), ref_perm)


@pytest.mark.parametrize("material,n_prim,pos_control", [(2, 1, False), (1, 2, False), (0, 1, False), (1, 1, True)])
def test_step_forward_parity(built_lib, material, n_prim, pos_control):
    conf = _conf(steps=8, n_primitive=n_prim, use_position_control=pos_control)
    B = 2
    sim = _sim(conf, B)
    kw = dict(v_scale=0.05, c_scale=1.0, f_scale=0.02, ylow=0.045) if material == 0 else {}  # liquid branch is centred on ylow
    st = util.mini_plasticine(sim, B, seed=material, material=material, **kw)
    act = _actions(B, n_prim).to(st.x.device)
    out, _ = sim.step_jax(st, act)
    ref = _oracle_step(conf, sim, st, act)
    ref64 = _oracle_step(conf, sim, st, act, torch.float64)
    for k in ("x", "v", "C", "F", "J"):
        e = util.rel_err(getattr(out, k), getattr(ref, k))
        floor = util.rel_err(getattr(ref, k), getattr(ref64, k))
        print(f"material={material} {k}: cuda-vs-oracle32 {e:.3e}   oracle32-vs-oracle64 {floor:.3e}")
        assert e < 1e-4, (k, e)
    for q in range(n_prim):
        for k in ("position", "rotation", "v", "w", "action_buffer"):
            e = util.rel_err(getattr(out.primitives[q], k), getattr(ref.primitives[q], k))
            assert e < 1e-5, (q, k, e)
    # mass is conserved by P2G/G2P bookkeeping: x moved by dt*v
    assert torch.isfinite(out.x).all()


def test_multi_step_episode_parity(built_lib):
    """6 env steps x 16 substeps from rest (F = I, v = 0), elasto-plastic block pushed by the box."""
    conf = _conf(steps=16)
    B = 2
    sim = _sim(conf, B)
    st = util.mini_plasticine(sim, B, seed=5, v_scale=0.0)
    n = st.x.shape[1]
    st = st._replace(C=torch.zeros_like(st.C), F=torch.eye(3, device=st.x.device).expand(B, n, 3, 3).contiguous())
    act = torch.tensor([[0.0, 0.0, 0.6, 0, 0, 0], [0.3, 0.0, 0.5, 0, 0, 0.2]], device=st.x.device)
    s_gpu, s_ref, s_ref64 = st, util.to_oracle_state(st), util.to_oracle_state(st, torch.float64)
    osim = omp.Simulator(util.oracle_conf(conf), sim.material.clone(), sim.h.clone())
    osim64 = omp.Simulator(util.oracle_conf(conf), sim.material.clone(), sim.h.clone().double(), dtype=torch.float64)
    for it in range(6):
        s_gpu, _ = sim.step_jax(s_gpu, act)
        s_ref = omp.step_batch(osim, s_ref, act.cpu())
        s_ref64 = omp.step_batch(osim64, s_ref64, act.cpu().double())
    for k in ("x", "v", "C", "F"):
        e = util.rel_err(getattr(s_gpu, k), getattr(s_ref, k))
        fl = util.rel_err(getattr(s_ref, k), getattr(s_ref64, k))   # fp32 noise of the reference arithmetic over the episode
        bar = util.floor_bar(1e-4, fl)
        print(f"episode {k}: cuda-vs-oracle32 {e:.3e}   oracle32-vs-oracle64 (floor) {fl:.3e}   bar {bar:.1e}")
        assert e < bar, (k, e, bar)                                   # north_star: state rtol 1e-4


def _leaf_list(state, n_prim):
    leaves = {k: getattr(state, k) for k in ("x", "v", "C", "F", "friction", "mu", "lamda")}
    for q in range(n_prim):
        for k in ("size", "friction", "position", "rotation", "action_scale"):
            leaves[f"p{q}.{k}"] = getattr(state.primitives[q], k)
    return leaves


def _run_grad(step_fn, state, action, cot, n_prim, to_dev):
    """L = sum(out_leaf * cot_leaf); returns d L / d (input leaves, action)."""
    leaves = _leaf_list(state, n_prim)
    req = {k: v.detach().clone().requires_grad_(True) for k, v in leaves.items()}
    prims = []
    for q, p in enumerate(state.primitives):
        if q < n_prim:
            p = p._replace(**{k.split(".")[1]: req[k] for k in req if k.startswith(f"p{q}.")})
        prims.append(p)
    st = state._replace(primitives=prims, **{k: req[k] for k in ("x", "v", "C", "F", "friction", "mu", "lamda")})
    a = action.detach().clone().requires_grad_(True)
    out = step_fn(st, a)
    L = 0
    for k in ("x", "v", "C", "F"):
        L = L + (getattr(out, k) * to_dev(cot[k])).sum()
    for q in range(n_prim):
        for k in ("position", "rotation"):
            L = L + (getattr(out.primitives[q], k) * to_dev(cot[f"p{q}.{k}"])).sum()
    names = list(req.keys())
    grads = torch.autograd.grad(L, [req[k] for k in names] + [a], allow_unused=True)
    res = {k: (g if g is not None else torch.zeros_like(req[k])) for k, g in zip(names, grads[:-1])}
    res["action"] = grads[-1]
    return res


@pytest.mark.parametrize("material,n_prim,pos_control,cot_scale",
                         [(2, 1, False, 1e-3), (1, 2, False, 1e-3), (0, 1, False, 1e-3), (1, 1, True, 1e-3),
                          (2, 1, False, 10.0), (-1, 1, False, 1e-3)])
def test_step_backward_parity(built_lib, material, n_prim, pos_control, cot_scale):
    """Adjoint of one step (4 substeps) against torch autograd of the oracle.  cot_scale=1e-3 keeps
    the per-env gradient norm < 1 (norm_grad is a pass-through), 10.0 exercises the renormalisation.
    material -1: liquid, elastic and plastic particles interleaved at random, so that nearly every warp mixes the
    SVD-free liquid path (UD_P2G_LIQUID_FAST) with the two SVD paths."""
    conf = _conf(steps=4, n_primitive=n_prim, use_position_control=pos_control)
    B = 2
    sim = _sim(conf, B)
    kw = dict(v_scale=0.05, c_scale=1.0, f_scale=0.02, ylow=0.045) if material == 0 else {}
    st = util.mini_plasticine(sim, B, seed=10 + material, material=2 if material < 0 else material, **kw)
    if material < 0:
        rs = np.random.RandomState(3)
        sim.material = torch.from_numpy(rs.randint(0, 3, sim.material.shape[0]).astype(np.int32))
        sim._material_dev = sim.material.to(st.x.device).contiguous()
        from unidom_b200 import _lib
        assert sim.params().p2g_mode & _lib.UD_P2G_LIQUID_FAST
    dev = st.x.device
    act = (_actions(B, n_prim, seed=4) * 0.8).to(dev)
    g = torch.Generator().manual_seed(7)
    S = conf.steps
    n = st.x.shape[1]
    cot = {"x": torch.randn((B, n, 3), generator=g), "v": torch.randn((B, n, 3), generator=g) * 0.1,
           "C": torch.randn((B, n, 3, 3), generator=g) * 1e-3, "F": torch.randn((B, n, 3, 3), generator=g) * 0.1}
    for q in range(n_prim):
        cot[f"p{q}.position"] = torch.randn((B, S, 3), generator=g)
        cot[f"p{q}.rotation"] = torch.randn((B, S, 4), generator=g)
    cot = {k: v * cot_scale for k, v in cot.items()}

    got = _run_grad(lambda s, a: sim.step_jax(s, a)[0], st, act, cot, n_prim, lambda t: t.to(dev))

    osim = omp.Simulator(util.oracle_conf(conf), sim.material.clone(), sim.h.clone())
    ost = util.to_oracle_state(st)
    ref = _run_grad(lambda s, a: omp.step_batch(osim, s, a), ost, act.cpu(), cot, n_prim, lambda t: t)
    osim64 = omp.Simulator(util.oracle_conf(conf), sim.material.clone(), sim.h.clone(), torch.float64)
    ref64 = _run_grad(lambda s, a: omp.step_batch(osim64, s, a), util.to_oracle_state(st, torch.float64),
                      act.cpu().double(), {k: v.double() for k, v in cot.items()}, n_prim, lambda t: t)
    worst = 0.0
    for k in ref:
        e = util.rel_err(got[k], ref[k])
        fl = util.rel_err(ref[k], ref64[k])
        cs = util.cosine(got[k], ref[k]) if float(ref[k].abs().max()) > 0 else 1.0
        print(f"grad {k:18s} rel {e:.3e} cos {cs:.6f}  | oracle32-vs-64 rel {fl:.3e}  max|ref| {float(ref[k].abs().max()):.3e}")
        if float(ref[k].abs().max()) > 1e-12:
            assert cs >= 0.999, (k, cs)
            assert e < max(1e-3, 20 * fl), (k, e, fl)


def test_p2g_deterministic_mode_is_bit_reproducible(built_lib):
    """UD_P2G_DETERMINISTIC: fixed-point integer accumulation across CTAs.  Two runs are bit-identical (the
    race canary of SURVEY section 5) and agree with the fp32-RED mode to rounding."""
    from unidom_b200 import _lib
    conf = _conf(steps=16)
    B = 3
    sim_d = _sim(conf, B, p2g_mode=_lib.UD_P2G_DETERMINISTIC)
    st = util.mini_plasticine(sim_d, B, seed=21, density=2.0)      # ~8 particles per cell: contended cells
    act = _actions(B, 1, seed=2).to(st.x.device)
    outs = []
    for _ in range(3):
        x = st.x.clone().requires_grad_(True)
        o, _ = sim_d.step_jax(st._replace(x=x), act)
        (gx,) = torch.autograd.grad((o.x * o.v).sum(), [x])
        outs.append((o, gx))
    for o, gx in outs[1:]:
        for k in ("x", "v", "C", "F", "J"):
            assert torch.equal(getattr(o, k), getattr(outs[0][0], k)), k
    sim_a = _sim(conf, B, p2g_mode=_lib.UD_P2G_ATOMIC)
    sim_a.material, sim_a.h = sim_d.material, sim_d.h
    sim_a.n_particles, sim_a._material_dev, sim_a._h_dev = sim_d.n_particles, sim_d._material_dev, sim_d._h_dev
    oa, _ = sim_a.step_jax(st, act)
    for k in ("x", "v", "C", "F"):
        e = util.rel_err(getattr(oa, k), getattr(outs[0][0], k))
        print(f"deterministic vs fp32-RED mode {k}: rel {e:.3e}")
        assert e < 2e-5, (k, e)
    ref = _oracle_step(conf, sim_d, st, act)
    assert util.rel_err(outs[0][0].x, ref.x) < 1e-4 and util.rel_err(outs[0][0].F, ref.F) < 1e-4


def test_taped_adjoint_matches_recompute_adjoint(built_lib):
    """ud_mpm_step_fwd_taped/_bwd_taped (residuals of every substep kept in HBM) against ud_mpm_step_fwd/_bwd (step
    input kept, substeps recomputed): the same kernels on the same data, so the forward is bit-identical in
    deterministic mode and the gradients differ only by the order of the fp32 REDs of G2P^T."""
    from unidom_b200 import _lib
    conf = _conf(steps=16, n_primitive=1)
    B = 3
    sims = {m: _sim(conf, B, p2g_mode=_lib.UD_P2G_DETERMINISTIC, adjoint=m) for m in ("tape", "recompute", "auto")}
    st = util.mini_plasticine(sims["tape"], B, seed=5, density=2.0)
    for m in ("recompute", "auto"):
        sims[m].material, sims[m].h, sims[m].n_particles = sims["tape"].material, sims["tape"].h, sims["tape"].n_particles
        sims[m]._material_dev, sims[m]._h_dev = sims["tape"]._material_dev, sims["tape"]._h_dev
    sims["auto"].tape_budget_bytes = 0                       # budget exhausted -> must fall back to recompute
    act = (_actions(B, 1, seed=6) * 0.8).to(st.x.device)
    g = torch.Generator().manual_seed(9)
    n = st.x.shape[1]
    cot = {"x": torch.randn((B, n, 3), generator=g) * 1e-3, "v": torch.randn((B, n, 3), generator=g) * 1e-4,
           "C": torch.randn((B, n, 3, 3), generator=g) * 1e-6, "F": torch.randn((B, n, 3, 3), generator=g) * 1e-4,
           "p0.position": torch.randn((B, 16, 3), generator=g) * 1e-3, "p0.rotation": torch.zeros((B, 16, 4))}
    res, outs, chose, held = {}, {}, {}, {}
    tape_bytes = sims["tape"]._L.ud_mpm_tape_bytes(__import__("ctypes").byref(sims["tape"].params()))
    for m, sim in sims.items():
        def step(s, a, sim=sim, m=m):
            outs[m] = sim.step_jax(s, a)[0]
            return outs[m]
        m0 = torch.cuda.memory_allocated()
        res[m] = _run_grad(step, st, act, cot, 1, lambda t: t.to(st.x.device))
        chose[m], held[m] = sim.last_adjoint, torch.cuda.memory_allocated() - m0
    assert chose == {"tape": "tape", "recompute": "recompute", "auto": "recompute"}, chose
    # the tape went back to the simulator's free list with the backward (outputs still held): the next step reuses it
    assert len(sims["tape"]._tape_pool) == 1 and sims["tape"]._tape_pool[0].numel() == tape_bytes + 256
    x2 = st.x.clone().requires_grad_(True)
    o2 = sims["tape"].step_jax(st._replace(x=x2), act)[0]
    assert len(sims["tape"]._tape_pool) == 0
    (o2.x.sum()).backward(retain_graph=True)
    assert len(sims["tape"]._tape_pool) == 1
    with pytest.raises(RuntimeError, match="tape was released"):
        (o2.v.sum()).backward()
    for k in ("x", "v", "C", "F", "J"):
        assert torch.equal(getattr(outs["tape"], k), getattr(outs["recompute"], k)), k
    assert torch.equal(outs["tape"].primitives[0].position, outs["recompute"].primitives[0].position)
    # the reverse pass is the same code on bit-identical residuals; what differs run to run is the order of the fp32
    # REDs of G2P^T, amplified by 16 substeps of SVD adjoints: "auto" IS a second recompute run = the noise floor
    for k in res["tape"]:
        if float(res["recompute"][k].abs().max()) <= 1e-20:
            continue
        e = util.rel_err(res["tape"][k], res["recompute"][k])
        fl = util.rel_err(res["auto"][k], res["recompute"][k])
        cs = util.cosine(res["tape"][k], res["recompute"][k])
        print(f"taped vs recompute grad {k:18s} rel {e:.3e} cos {cs:.8f} | recompute vs recompute rel {fl:.3e}")
        assert cs > 0.999999 and e < max(1e-3, 10 * fl), (k, e, fl, cs)


def test_edge_cases_out_of_grid_nan_single_env(built_lib):
    """B = 1, a particle count that is not a multiple of the CTA size, particles outside the grid (scatter
    dropped / gather clamped / negative index wrap), NaN in the input state (norm_grad_state's
    forward nan_to_num, mpm_simulator.py:376-381): forward and adjoint against the oracle."""
    conf = _conf(steps=3)
    B = 1
    sim = _sim(conf, B)
    st = util.mini_plasticine(sim, B, seed=31, material=1)
    n = st.x.shape[1]
    assert n % 64 != 0
    x, v, F = st.x.clone(), st.v.clone(), st.F.clone()
    x[0, 0] = torch.tensor([-0.004, 0.03, 0.25], device=x.device)       # base x = -0 / negative side: wraps
    x[0, 1] = torch.tensor([0.25, 0.03, 0.4995], device=x.device)       # touches the +z face of res (48/96 = 0.5)
    x[0, 2] = torch.tensor([0.25, 0.34, 0.25], device=x.device)         # above the grid (res_y = 32/96 = 0.333)
    v[0, 3, 1] = float("nan")
    F[0, 4, 0, 1] = float("nan")
    st = st._replace(x=x, v=v, F=F)
    act = _actions(B, 1, seed=9).to(x.device)
    g = torch.Generator().manual_seed(5)
    cot = {"x": torch.randn((B, n, 3), generator=g) * 1e-3, "v": torch.randn((B, n, 3), generator=g) * 1e-4,
           "C": torch.zeros((B, n, 3, 3)), "F": torch.zeros((B, n, 3, 3)),
           "p0.position": torch.zeros((B, conf.steps, 3)), "p0.rotation": torch.zeros((B, conf.steps, 4))}
    got = _run_grad(lambda s, a: sim.step_jax(s, a)[0], st, act, cot, 1, lambda t: t.to(x.device))
    osim = omp.Simulator(util.oracle_conf(conf), sim.material.clone(), sim.h.clone())
    ost = util.to_oracle_state(st)
    ref = _run_grad(lambda s, a: omp.step_batch(osim, s, a), ost, act.cpu(), cot, 1, lambda t: t)
    out, _ = sim.step_jax(st, act)
    oref = omp.step_batch(osim, ost, act.cpu())
    mask = torch.ones(n, dtype=torch.bool)
    mask[0] = False     # the wrapped particle is alone on the far face: v = p / m with m ~ 1e-12 (compared loosely)
    assert util.rel_err(out.v[0, 0], oref.v[0, 0]) < 5e-2 and util.rel_err(out.x[0, 0], oref.x[0, 0]) < 1e-3
    for k in ("x", "v", "C"):
        a, b = getattr(out, k)[0].cpu()[mask], getattr(oref, k)[0][mask]
        assert torch.isfinite(a).all(), k
        assert util.rel_err(a, b) < 1e-4, (k, util.rel_err(a, b))
    for k in ("x", "v", "action"):
        a, b = got[k].cpu(), ref[k]
        if k != "action":
            a, b = a[0][mask], b[0][mask]
        assert util.cosine(a, b) > 0.999, (k, util.cosine(a, b))


def test_invalid_arguments_are_rejected(built_lib):
    """Error behaviour of the C ABI: nothing is enqueued, a negative status comes back (no exception in C)."""
    import ctypes as C
    from unidom_b200 import _lib
    conf = _conf(steps=2)
    sim = _sim(conf, 2)
    st = util.mini_plasticine(sim, 2, seed=1)
    L = _lib.lib()
    p = sim.params(B=2, n=st.x.shape[1])
    assert L.ud_mpm_fwd_workspace_bytes(C.byref(p)) > 0
    p.p2g_mode = 7
    assert L.ud_mpm_fwd_workspace_bytes(C.byref(p)) == 0
    with pytest.raises(RuntimeError):
        sim.p2g_mode = 7
        sim.step_jax(st, _actions(2, 1).to(st.x.device))


def test_cuda_graph_replay_of_the_forward_scan(built_lib):
    """unidom_b200.graphs.GraphedMPMScan: T step_jax calls captured once, replayed with new inputs; deterministic P2G
    makes replay and eager launches bit-identical."""
    from unidom_b200 import _lib
    from unidom_b200.graphs import GraphedMPMScan
    conf = _conf(steps=8)
    B, T = 2, 3
    sim = _sim(conf, B, p2g_mode=_lib.UD_P2G_DETERMINISTIC)
    st = util.mini_plasticine(sim, B, seed=41)
    acts = torch.stack([_actions(B, 1, seed=50 + t) for t in range(T)]).to(st.x.device)
    graph = GraphedMPMScan(sim, st, acts)
    for trial in range(2):                                   # second trial: different inputs through the same graph
        s_in = st._replace(x=st.x + 1e-3 * trial, v=st.v * (1 + trial))
        a_in = acts * (1 - 0.3 * trial)
        out = graph(s_in, a_in)
        ref = s_in
        with torch.no_grad():
            for t in range(T):
                ref, _ = sim.step_jax(ref, a_in[t])
        for k in ("x", "v", "C", "F", "J"):
            assert torch.equal(getattr(out, k), getattr(ref, k)), (trial, k)
        assert torch.equal(out.primitives[0].position, ref.primitives[0].position)


@pytest.mark.parametrize("window", [1, 3, 4, 8])
def test_windowed_adjoint_matches_full_recompute(built_lib, window):
    """ud_mpm_step_bwd_windowed (K-spaced substep checkpoints inside the step, north_star (e)) against ud_mpm_step_bwd:
    the same kernels on the same data -- with deterministic P2G the recomputed trajectory is bit-identical, so the
    gradients differ only by the order of the fp32 REDs of G2P^T (floor: a second full-recompute run).  K = 3 does
    not divide S = 10 (ragged last window); K = 1 keeps every substep as a checkpoint; the workspace shrinks with K."""
    from unidom_b200 import _lib
    import ctypes as C
    conf = _conf(steps=10, n_primitive=1)
    B = 2
    sims = {m: _sim(conf, B, p2g_mode=_lib.UD_P2G_DETERMINISTIC, ckpt_window=w)
            for m, w in (("full", None), ("full2", None), ("win", window))}
    st = util.mini_plasticine(sims["full"], B, seed=15, density=2.0)
    for m in ("full2", "win"):
        sims[m].material, sims[m].h, sims[m].n_particles = sims["full"].material, sims["full"].h, sims["full"].n_particles
        sims[m]._material_dev, sims[m]._h_dev = sims["full"]._material_dev, sims["full"]._h_dev
    act = (_actions(B, 1, seed=16) * 0.8).to(st.x.device)
    g = torch.Generator().manual_seed(19)
    n = st.x.shape[1]
    cot = {"x": torch.randn((B, n, 3), generator=g) * 1e-3, "v": torch.randn((B, n, 3), generator=g) * 1e-4,
           "C": torch.randn((B, n, 3, 3), generator=g) * 1e-6, "F": torch.randn((B, n, 3, 3), generator=g) * 1e-4,
           "p0.position": torch.randn((B, 10, 3), generator=g) * 1e-3, "p0.rotation": torch.zeros((B, 10, 4))}
    res = {m: _run_grad(lambda s, a, sim=sim: sim.step_jax(s, a)[0], st, act, cot, 1, lambda t: t.to(st.x.device))
           for m, sim in sims.items()}
    for k in res["full"]:
        ref = res["full"][k]
        if float(ref.abs().max()) == 0.0:
            assert float(res["win"][k].abs().max()) == 0.0, k
            continue
        floor = util.rel_err(res["full2"][k], ref)
        e = util.rel_err(res["win"][k], ref)
        print(f"windowed K={window} grad {k:12s}: rel {e:.3e}  floor (two full-recompute runs) {floor:.3e}")
        assert e <= max(1e-5, 20 * floor) and util.cosine(res["win"][k], ref) > 0.999999, (k, e, floor)
    p = sims["win"].params(B=B, n=n)
    full_b = sims["win"]._L.ud_mpm_bwd_workspace_bytes(C.byref(p))
    win_b = sims["win"]._L.ud_mpm_bwd_windowed_workspace_bytes(C.byref(p), window)
    print(f"windowed K={window}: workspace {win_b / 1e6:.2f} MB vs {full_b / 1e6:.2f} MB store-all-substeps")
    if window <= 4:
        assert win_b < full_b


def test_cuda_graph_of_the_differentiated_scan(built_lib):
    """graphs.GraphedMPMScanGrad: T step_jax calls and their adjoints as two CUDA graphs behind one autograd Function,
    against the eager scan (deterministic P2G: forward bit-identical; gradients equal up to the order of the fp32 REDs
    of G2P^T, floor = a second eager run); replayed with new inputs."""
    from unidom_b200 import _lib
    from unidom_b200.graphs import GraphedMPMScanGrad
    conf = _conf(steps=6)
    B, T = 2, 3
    sim = _sim(conf, B, p2g_mode=_lib.UD_P2G_DETERMINISTIC)
    st = util.mini_plasticine(sim, B, seed=43)
    acts = torch.stack([_actions(B, 1, seed=60 + t) for t in range(T)]).to(st.x.device) * 0.6
    scan = GraphedMPMScanGrad(sim, st, acts)
    w = torch.linspace(-1, 1, st.x.numel(), device=st.x.device).reshape(st.x.shape) * 1e-3

    def eager(x, a):
        s = st._replace(x=x)
        for t in range(T):
            s, _ = sim.step_jax(s, a[t])
        return s
    for trial in range(2):
        x0 = st.x + 1e-3 * trial
        a0 = acts * (1 - 0.2 * trial)
        res = []
        for fn in (lambda x, a: scan(st._replace(x=x), a), eager, eager):
            x = x0.clone().requires_grad_(True)
            a = a0.clone().requires_grad_(True)
            out = fn(x, a)
            gx, ga = torch.autograd.grad((out.x * w).sum() + (out.v * w).sum() * 0.1, [x, a])
            res.append((out, gx, ga))
        for k in ("x", "v", "C", "F", "J"):
            assert torch.equal(getattr(res[0][0], k), getattr(res[1][0], k)), (trial, k)
        for i, name in ((1, "x"), (2, "action")):
            e, fl = util.rel_err(res[0][i], res[1][i]), util.rel_err(res[2][i], res[1][i])
            print(f"graphed vs eager scan, trial {trial}: grad {name} rel {e:.3e} (eager vs eager {fl:.3e})")
            assert e <= max(1e-5, 20 * fl), (name, e, fl)
    scan.close()
