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
