"""Test helpers: move states between the product (cuda torch, batched) and the oracle (cpu torch)."""
import numpy as np
import torch

from oracle import mpm as omp
from oracle import primitives as oP


def oracle_conf(conf):
    return omp.MPMConf(n_grid=conf.n_grid, res=tuple(conf.res), dt=conf.dt, steps=conf.steps, E=conf.E, nu=conf.nu,
                       ground_friction=conf.ground_friction, gravity=tuple(conf.gravity),
                       n_primitive=conf.n_primitive, sdf_kind=conf.sdf_kind,
                       use_position_control=conf.use_position_control, p_rho=conf.p_rho)


def to_oracle_state(state, dtype=torch.float32):
    """product MPMState (batched, any device) -> oracle MPMState (batched, cpu)."""
    def cv(t):
        t = t.detach().cpu()
        return t.to(dtype) if t.is_floating_point() else t
    prims = [oP.PrimitiveState(*[cv(t) for t in p]) for p in state.primitives]
    vals = {k: cv(getattr(state, k)) for k in state._fields if k != "primitives"}
    return omp.MPMState(primitives=prims, **vals)


def rel_err(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def cosine(a, b):
    a = a.detach().cpu().double().flatten()
    b = b.detach().cpu().double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def mini_plasticine(sim, B, seed=0, density=1.0, v_scale=0.3, material=2, ylow=0.02, c_scale=2.0, f_scale=0.05):
    """A small block near the ground with a box primitive cutting its edge, random velocities,
    slightly non-identity F and non-zero C so that every term of the substep is exercised."""
    from unidom_b200.mpm_simulator import create_primitive
    conf = sim.conf
    g = torch.Generator().manual_seed(seed)
    state = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.1, 0.06, 0.08], init_pos=[0.25, ylow, 0.25],
                        z_rotation_angle=0, material=material, density=density)
    state.primitives.append(create_primitive(conf, friction=0.9, softness=666, color=[0.5, 0.5, 0.5],
                                             size=[0.015, 0.06, 0.015], init_pos=[0.25, 0.01, 0.205]))
    for _ in range(1, conf.n_primitive):
        state.primitives.append(create_primitive(conf, friction=0.5, softness=666, color=[0.5, 0.5, 0.5],
                                                 size=[0.02, 0.03, 0.01], init_pos=[0.29, 0.02, 0.27]))
    # realistic primitive speeds: |action| <= 1 moves the tool by <= 0.012 per step
    state = state._replace(primitives=[p._replace(action_scale=p.action_scale * 0.012) for p in state.primitives])
    st = sim.reset_jax(state)
    n = st.x.shape[1]
    dev = st.x.device
    v = (torch.randn((B, n, 3), generator=g) * v_scale).to(dev)
    Cm = (torch.randn((B, n, 3, 3), generator=g) * c_scale).to(dev)
    F = (torch.eye(3)[None, None] + f_scale * torch.randn((B, n, 3, 3), generator=g)).to(dev)
    x = st.x + (torch.randn((B, n, 3), generator=g) * 1e-3).to(dev)
    return st._replace(x=x, v=v, C=Cm, F=F)


# ---------------------------------------------------------------------------------------------------
# Golden fixtures produced by oracle/gen_golden.py (the unmodified reference sources under oracle/jaxshim)
# ---------------------------------------------------------------------------------------------------
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STATE_F = ("x", "v", "C", "F", "J", "friction", "mu", "lamda")
PRIM_F = ("size", "friction", "softness", "position", "rotation", "v", "w", "action_buffer", "action_scale")


def golden_mpm(name):
    """-> (conf for the product, dict of torch tensors) from tests/golden/ref_mpm_<name>.npz."""
    from unidom_b200 import confs
    d = {k: v for k, v in np.load(os.path.join(GOLD, f"ref_mpm_{name}.npz")).items()}
    n_grid, rx, ry, rz, steps, n_prim, pos_control, sdf_kind = [int(v) for v in d["conf"]]
    dt, E, nu, gf, g0, g1, g2 = [float(v) for v in d["conf_f"]]
    conf = confs.MPMConf(n_grid=n_grid, res=(rx, ry, rz), dt=dt, steps=steps, E=E, nu=nu, ground_friction=gf,
                         gravity=(g0, g1, g2), n_primitive=n_prim, sdf_kind=sdf_kind,
                         use_position_control=bool(pos_control))
    return conf, {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in d.items()}


def golden_oracle_state(conf, d, prefix="in_", dtype=torch.float32):
    """oracle MPMState (batched) from a golden dict."""
    B = d[prefix + "x"].shape[0]
    prims = []
    for q in range(conf.n_primitive):
        f = {k: d[f"{prefix}p{q}_{k}"].to(dtype) for k in PRIM_F}
        prims.append(oP.PrimitiveState(
            size=f["size"], dim=torch.full((B, 1), 3, dtype=torch.int32), friction=f["friction"], softness=f["softness"],
            color=torch.full((B, 3), 0.5, dtype=dtype), position=f["position"], rotation=f["rotation"], v=f["v"], w=f["w"],
            xyz_limit=torch.tensor([[0.0, 1.0]] * 3, dtype=dtype)[None].repeat(B, 1, 1), action_buffer=f["action_buffer"],
            action_scale=f["action_scale"], min_dist=torch.zeros(B, dtype=torch.int32),
            dist_norm=torch.zeros(B, dtype=torch.int32)))
    vals = {k: d[prefix + k].to(dtype) for k in STATE_F}
    return omp.MPMState(cur_step=torch.zeros(B, dtype=torch.int32), primitives=prims,
                        key=torch.zeros((B, 2), dtype=torch.int32), **vals)


def golden_product_state(sim, conf, d, device="cuda"):
    """product MPMState on `device` from a golden dict; also sets sim.material / sim.h."""
    from unidom_b200.mpm_simulator import MPMState, PrimitiveState
    ost = golden_oracle_state(conf, d)
    sim.material, sim.h = d["material"].to(torch.int32), d["h"].to(torch.float32)
    sim.n_particles = ost.x.shape[1]
    sim._material_dev = sim.material.to(device).contiguous()
    sim._h_dev = sim.h.to(device).contiguous()
    prims = [PrimitiveState(*[t.to(device) for t in p]) for p in ost.primitives]
    vals = {k: getattr(ost, k).to(device) for k in ost._fields if k != "primitives"}
    return MPMState(primitives=prims, **vals)


MPM_GRAD_LEAVES = ("x", "v", "C", "F", "friction", "mu", "lamda")
MPM_PRIM_GRAD_LEAVES = ("size", "friction", "position", "rotation", "action_scale")


def golden_mpm_grads(step_fn, state, action, d, n_prim, to_dev):
    """d L / d (input leaves, action) for L = sum(out_leaf * cot_leaf) with the golden cotangents."""
    req = {k: getattr(state, k).detach().clone().requires_grad_(True) for k in MPM_GRAD_LEAVES}
    prims = []
    for q, p in enumerate(state.primitives):
        if q < n_prim:
            pr = {k: getattr(p, k).detach().clone().requires_grad_(True) for k in MPM_PRIM_GRAD_LEAVES}
            req.update({f"p{q}_{k}": v for k, v in pr.items()})
            p = p._replace(**pr)
        prims.append(p)
    st = state._replace(primitives=prims, **{k: req[k] for k in MPM_GRAD_LEAVES})
    a = action.detach().clone().requires_grad_(True)
    out = step_fn(st, a)
    L = 0
    for k in ("x", "v", "C", "F"):
        L = L + (getattr(out, k) * to_dev(d["cot_" + k]).to(getattr(out, k).dtype)).sum()
    for q in range(n_prim):
        for k in ("position", "rotation"):
            t = getattr(out.primitives[q], k)
            L = L + (t * to_dev(d[f"cot_p{q}_{k}"]).to(t.dtype)).sum()
    names = list(req.keys())
    grads = torch.autograd.grad(L, [req[k] for k in names] + [a], allow_unused=True)
    res = {k: (g if g is not None else torch.zeros_like(req[k])) for k, g in zip(names, grads[:-1])}
    res["action"] = grads[-1]
    return res, out


_FLOORS = {}


def golden_mpm_floor(name):
    """fp32 noise floor of the REFERENCE ARITHMETIC on golden case `name`: the oracle restatement (pinned to the reference
    sources by tests/test_oracle_golden.py) run in fp32 and in fp64 from the fixture's inputs and cotangents.
    -> ({state leaf: rel}, {gradient leaf: rel}); a tolerance above north_star's bar is only ever asserted as a multiple
    of these measured numbers."""
    if name in _FLOORS:
        return _FLOORS[name]
    conf, d = golden_mpm(name)
    res = {}
    for dtype in (torch.float32, torch.float64):
        osim = omp.Simulator(oracle_conf(conf), d["material"].to(torch.int32), d["h"].to(dtype), dtype=dtype)
        st = golden_oracle_state(conf, d, dtype=dtype)
        act = d["action"].to(dtype)
        if "g_x" in d:
            got, out = golden_mpm_grads(lambda s, a: omp.step_batch(osim, s, a), st, act, d, conf.n_primitive, lambda t: t)
        else:
            with torch.no_grad():
                got, out = {}, omp.step_batch(osim, st, act)
        res[dtype] = (out, got)
    (o32, g32), (o64, g64) = res[torch.float32], res[torch.float64]
    fs = {k: rel_err(getattr(o32, k), getattr(o64, k)) for k in ("x", "v", "C", "F", "J")}
    fg = {k: rel_err(g32[k], g64[k]) for k in g32 if g32[k] is not None and float(g64[k].abs().max()) > 1e-20}
    _FLOORS[name] = (fs, fg)
    return fs, fg


def env_floor(name):
    """fp32 noise of the UNMODIFIED reference on env fixture `name`: tests/golden/ref_mpmenv_<name>_f64.npz holds the same
    rollout of the same reference env classes, from the fixture's own inputs, with the shim's float type switched to
    float64 (oracle/gen_golden.py --f64).  -> f(key) = rel_err(fp32 fixture[key], fp64 run[key])."""
    a = np.load(os.path.join(GOLD, f"ref_mpmenv_{name}.npz"))
    b = np.load(os.path.join(GOLD, f"ref_mpmenv_{name}_f64.npz"))

    def floor(key):
        return rel_err(torch.from_numpy(a[key]).double(), torch.from_numpy(b[key]).double())
    return floor


K_FLOOR = 4.0      # bars above north_star's are K_FLOOR x a measured floor, never a hand-picked constant


def floor_bar(bar, floor):
    return max(bar, K_FLOOR * floor)
