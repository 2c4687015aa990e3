"""MLS-MPM simulator step -- torch-CPU restatement (test infrastructure).

Follows core/engine/mpm_simulator.py (line numbers cited per function).
Written per environment (the reference applies jax.vmap over axis 0,
mpm_simulator.py:61-63); ``step_batch`` loops over the batch.

Reverse mode: torch autograd of this forward.  The reference's custom VJPs are
mirrored where they change the numbers:
  * substep_wrapper (:332-363) recomputes the substep and differentiates
    sum(nstate * g) over x, v, C, F, friction, mu, lamda and the primitive
    leaves only -> the cotangent of J' is dropped (J' is detached here);
  * norm_grad_state / norm_grad (:375-411) -> ``NormGrad`` below.
JAX semantics mirrored explicitly: astype(int32) truncation, negative index
wrap, out-of-bounds scatter DROP / gather CLAMP, array.trace() over axes (0,1).
"""
from dataclasses import dataclass, field
from typing import List, NamedTuple, Tuple

import torch

from . import primitives as P
from .svd import svd


class MPMState(NamedTuple):  # mpm_simulator.py:13-24 (same field order)
    x: torch.Tensor = None
    v: torch.Tensor = None
    C: torch.Tensor = None
    F: torch.Tensor = None
    J: torch.Tensor = None
    cur_step: torch.Tensor = None
    primitives: List[P.PrimitiveState] = []
    key: torch.Tensor = None
    friction: torch.Tensor = None
    mu: torch.Tensor = None
    lamda: torch.Tensor = None


@dataclass
class MPMConf:
    """Scalar configuration (the reference's per-task DefaultConf classes)."""
    n_grid: int = 64
    res: Tuple[int, int, int] = (32, 32, 32)
    dt: float = 1e-4
    steps: int = 16
    E: float = 100.0
    nu: float = 0.1
    ground_friction: float = 0.1
    gravity: Tuple[float, float, float] = (0.0, -9.8, 0.0)
    n_primitive: int = 1
    sdf_kind: int = P.SDF_BOX
    use_position_control: bool = False
    p_rho: float = 1.0

    @property
    def dx(self):
        return 1 / self.n_grid

    @property
    def inv_dx(self):
        return float(self.n_grid)

    @property
    def p_vol(self):
        return (self.dx * 0.5) ** 2

    @property
    def p_mass(self):
        return self.p_vol * self.p_rho


def _offsets():
    """mpm_simulator.py:50-51: the 27 stencil offsets, i-major."""
    a, b, c = torch.meshgrid(torch.arange(3), torch.arange(3), torch.arange(3), indexing="ij")
    return torch.stack([a, b, c], dim=-1).reshape(-1, 3)


class NormGrad(torch.autograd.Function):
    """norm_grad_state / norm_grad (mpm_simulator.py:375-411).

    forward: identity (first ``n_scrub`` leaves get nan_to_num, :376-381);
    backward: nan_to_num every cotangent, global L2 norm over all leaves,
    divide by it unless it is < 1 (:389-394, :403-408).
    """

    @staticmethod
    def forward(ctx, n_scrub, *leaves):
        out = []
        for i, t in enumerate(leaves):
            out.append(torch.nan_to_num(t) if i < n_scrub else t.clone())
        return tuple(out)

    @staticmethod
    def backward(ctx, *gs):
        gs = [torch.nan_to_num(g + 0.0) for g in gs]
        g_norm = torch.sqrt(sum((g * g).sum() for g in gs))
        if not bool(g_norm < 1.0):
            gs = [g / g_norm for g in gs]
        return (None, *gs)


_PRIM_FLOAT_FIELDS = ("size", "friction", "softness", "color", "position", "rotation", "v", "w",
                      "xyz_limit", "action_buffer", "action_scale")


class Simulator:
    """Per-environment functional core of SimpleMPMSimulator."""

    def __init__(self, conf: MPMConf, material: torch.Tensor, h: torch.Tensor, dtype=torch.float32,
                 checkpoint_substeps=False):
        # checkpoint_substeps: torch.utils.checkpoint around every substep (same numbers; autograd keeps one substep's
        # graph at a time instead of S of them: full-size scenes with S = 70 would otherwise need tens of GB on the host)
        self.checkpoint_substeps = checkpoint_substeps
        self.conf = conf
        self.material = material
        self.h = h.to(dtype)
        self.dtype = dtype
        self.idx = _offsets()
        res = conf.res
        a, b, c = torch.meshgrid(torch.arange(res[0]), torch.arange(res[1]), torch.arange(res[2]), indexing="ij")
        self.grid_idx_3d = torch.stack([a, b, c], dim=-1)            # (X,Y,Z,3) int64
        self.gravity = torch.tensor(conf.gravity, dtype=dtype)

    # ---------------------------------------------------------------- substep
    def _stencil(self, x):
        """mpm_simulator.py:233-235."""
        c = self.conf
        base = (x * c.inv_dx - 0.5).to(torch.int32)
        fx = x * c.inv_dx - base.to(x.dtype)
        w = torch.stack([0.5 * (1.5 - fx) ** 2, 0.75 - (fx - 1) ** 2, 0.5 * (fx - 0.5) ** 2])
        return base, fx, w

    def _lin_index(self, pos, clamp):
        """Linear cell index under JAX scatter/gather rules (negative wraps, then drop/clamp)."""
        res = torch.tensor(self.conf.res, dtype=torch.int64)
        pos = pos.to(torch.int64)
        pos = torch.where(pos < 0, pos + res, pos)
        valid = ((pos >= 0) & (pos < res)).all(-1)
        if clamp:
            pos = torch.minimum(torch.maximum(pos, torch.zeros_like(res)), res - 1)
        lin = (pos[..., 0] * res[1] + pos[..., 1]) * res[2] + pos[..., 2]
        return lin, valid

    def constitutive(self, state):
        """mpm_simulator.py:238-268.  Returns (F2, affine)."""
        c = self.conf
        dt, dtype = c.dt, state.x.dtype
        eye = torch.eye(3, dtype=dtype)
        liquid = self.material == 0
        plastic = self.material == 2
        F_ = (eye[None] + dt * state.C) @ state.F
        h = self.h.to(dtype).clamp(0.1, 5)
        mu, la = state.mu * h, state.lamda * h
        mu = torch.where(liquid, torch.zeros((), dtype=dtype), mu)
        la = torch.where(liquid, torch.ones((), dtype=dtype), la)
        U, sig, V = svd(F_)
        sig_c = P.jclip(sig, 1 - 2.5e-2 * 10, 1 + 4.5e-3 * 100)
        sig = torch.where(plastic[:, None], sig_c, sig)
        J = sig.prod(-1)[:, None, None]
        sig_m = eye[None] * sig[..., None]
        F2 = torch.where(plastic[:, None, None], U @ sig_m @ V, F_)
        stress = 2 * mu[:, None, None] * (F2 - U @ V) @ F2.transpose(1, 2) \
            + eye[None] * la[:, None, None] * J * (J - 1)
        stress = (-dt * c.p_vol * 4) * stress / c.dx ** 2
        affine = stress + c.p_mass * state.C
        return F2, affine

    def p2g(self, v, fx, w, base, affine):
        """p2g_micro, mpm_simulator.py:178-194."""
        c = self.conf
        n = v.shape[0]
        idx = self.idx
        offset = idx[:, None, :].expand(27, n, 3)
        dpos = (offset.to(v.dtype) - fx[None]) * c.dx
        weight = w[idx[:, 0]][:, :, 0] * w[idx[:, 1]][:, :, 1] * w[idx[:, 2]][:, :, 2]     # (27,n)
        pos = base[None].to(torch.int64) + offset
        vals = weight[..., None] * (c.p_mass * v[None] + (affine[None] @ dpos[..., None]).squeeze(-1))
        lin, valid = self._lin_index(pos.reshape(-1, 3), clamp=False)
        G = c.res[0] * c.res[1] * c.res[2]
        grid_m = torch.zeros(G, dtype=v.dtype).index_add(0, lin[valid], (weight.flatten() * c.p_mass)[valid])
        grid_v = torch.zeros(G, 3, dtype=v.dtype).index_add(0, lin[valid], vals.reshape(-1, 3)[valid])
        return grid_v.reshape(c.res + (3,)), grid_m.reshape(c.res)

    def grid_op(self, f, grid_v, grid_m, state):
        """mpm_simulator.py:283-313.  state.primitives already advanced by FK."""
        c = self.conf
        dtype = grid_v.dtype
        grid_v_ = grid_v / grid_m[..., None]
        grid_v = torch.where(grid_m[..., None] > 0, grid_v_, grid_v)
        grid_v = grid_v + c.dt * self.gravity.to(dtype)
        gi = self.grid_idx_3d
        grid_pos = gi.to(dtype) * c.dx
        for i in range(c.n_primitive):
            if c.use_position_control:
                grid_v = P.position_control(f, grid_pos, grid_v, c.dt, state.primitives[i], c.sdf_kind)
            else:
                grid_v = P.collide(f, grid_pos, grid_v, c.dt, state.primitives[i], c.sdf_kind)
        # ground friction (:297-307)
        normal = torch.tensor([0.0, 1.0, 0.0], dtype=dtype)
        lin = grid_v[..., 1] + 1e-30
        gi_eps = gi.to(dtype) * 1e-30
        vit = grid_v - lin[..., None] * normal.reshape(1, 1, 1, 3) - gi_eps
        lit = torch.sqrt(((vit + 1e-12) ** 2).sum(-1))
        grid_v_ = P.jclip(1.0 + state.friction * lin[..., None] / lit[..., None], 0.0) * (vit + gi_eps)
        grid_v_ = torch.cat([grid_v_[..., 0:1], torch.zeros_like(grid_v_[..., 1:2]), grid_v_[..., 2:3]], dim=-1)
        friction_mask = (gi[..., 1] < 3)
        fric_speed_mask = grid_v[..., 1] <= 0
        grid_v = torch.where((friction_mask & fric_speed_mask)[..., None], grid_v_, grid_v)
        # boundary (:310-313) -- note n_grid, not res
        cond = ((gi < 3) & (grid_v < 0)) | ((gi > c.n_grid - 3) & (grid_v > 0))
        grid_v = torch.where(cond, torch.zeros((), dtype=dtype), grid_v)
        return grid_v

    def g2p(self, grid_v, fx, w, base):
        """g2p_micro, mpm_simulator.py:196-221."""
        c = self.conf
        n = fx.shape[0]
        idx = self.idx
        offset = idx[:, None, :].expand(27, n, 3)
        dpos = offset.to(fx.dtype) - fx[None]
        weight = w[idx[:, 0]][:, :, 0] * w[idx[:, 1]][:, :, 1] * w[idx[:, 2]][:, :, 2]
        pos = base[None].to(torch.int64) + offset
        lin, _ = self._lin_index(pos.reshape(-1, 3), clamp=True)
        g_v = grid_v.reshape(-1, 3)[lin].reshape(27, n, 3)
        new_v = (weight[..., None] * g_v).sum(0)
        outer = g_v[..., :, None] * dpos[..., None, :]
        new_C = (4 * weight[..., None, None] * outer * c.inv_dx).sum(0)
        return new_v, new_C

    def substep(self, f, state: MPMState) -> MPMState:
        """mpm_simulator.py:223-330."""
        c = self.conf
        base, fx, w = self._stencil(state.x)
        F2, affine = self.constitutive(state)
        grid_v, grid_m = self.p2g(state.v, fx, w, base, affine)
        prims = [P.forward_kinematics(f, p) for p in state.primitives]     # :277-278
        state = state._replace(F=F2, primitives=prims)
        grid_v = self.grid_op(f, grid_v, grid_m, state)
        v_, C_ = self.g2p(grid_v, fx, w, base)
        x_ = state.x + c.dt * v_
        k = min(3, C_.shape[0])
        tr = sum(C_[i, i, :] for i in range(k)).sum(-1)                      # :327 (trace over axes 0,1)
        J_ = (state.J * (1 + c.dt * tr)).detach()                           # cotangent dropped (:343-350)
        return state._replace(x=x_, v=v_, C=C_, J=J_)

    # ------------------------------------------------------------------- step
    def step(self, state: MPMState, action: torch.Tensor) -> MPMState:
        """mpm_simulator.py:413-429 (returns the carry; the reference returns it twice)."""
        c = self.conf
        state, action = self._norm_grad_in(state, action)
        action = P.jclip(action, -1, 1)
        prims = [P.set_action(c.steps, action[i * 6:(i + 1) * 6], state.primitives[i])
                 for i in range(c.n_primitive)] + list(state.primitives[c.n_primitive:])
        state = state._replace(primitives=prims)
        for f in range(c.steps):
            state = self._substep_ckpt(f, state) if self.checkpoint_substeps else self.substep(f, state)
        # copy_frame(steps, 0) (:365-373): source row S clamps to S-1
        prims = []
        for i, p in enumerate(state.primitives):
            if i < c.n_primitive:
                position = torch.cat([P._row(p.position, c.steps)[None], p.position[1:]], dim=0)
                rotation = torch.cat([P._row(p.rotation, c.steps)[None], p.rotation[1:]], dim=0)
                p = p._replace(position=position, rotation=rotation)
            prims.append(p)
        return state._replace(primitives=prims)

    def _substep_ckpt(self, f, state):
        from torch.utils.checkpoint import checkpoint
        n_prim = len(state.primitives)
        fields = [k for k in state._fields if k != "primitives"]
        pf = P.PrimitiveState._fields

        def pack(st):
            return [getattr(st, k) for k in fields] + [getattr(p, k) for p in st.primitives for k in pf]

        def unpack(ls):
            vals = dict(zip(fields, ls[:len(fields)]))
            rest = ls[len(fields):]
            prims = [P.PrimitiveState(*rest[i * len(pf):(i + 1) * len(pf)]) for i in range(n_prim)]
            return MPMState(primitives=prims, **vals)

        out = checkpoint(lambda *ls: tuple(pack(self.substep(f, unpack(list(ls))))), *pack(state), use_reentrant=False)
        return unpack(list(out))

    def _norm_grad_in(self, state, action):
        """norm_grad_state(state), norm_grad(action) (:415-416)."""
        leaves = [state.x, state.v, state.C, state.F, state.J, state.friction, state.mu, state.lamda]
        for p in state.primitives:
            leaves += [getattr(p, k) for k in _PRIM_FLOAT_FIELDS]
        out = list(NormGrad.apply(5, *leaves))
        x, v, C, F, J, fr, mu, la = out[:8]
        rest = out[8:]
        prims = []
        nf = len(_PRIM_FLOAT_FIELDS)
        for i, p in enumerate(state.primitives):
            vals = rest[i * nf:(i + 1) * nf]
            prims.append(p._replace(**dict(zip(_PRIM_FLOAT_FIELDS, vals))))
        state = state._replace(x=x, v=v, C=C, F=F, J=J, friction=fr, mu=mu, lamda=la, primitives=prims)
        (action,) = NormGrad.apply(0, action)
        return state, action


# ------------------------------------------------------------------ batching
def tree_map_state(fn, state: MPMState) -> MPMState:
    prims = [P.PrimitiveState(*[fn(t) for t in p]) for p in state.primitives]
    vals = {k: (fn(getattr(state, k)) if getattr(state, k) is not None else None)
            for k in state._fields if k != "primitives"}
    return MPMState(primitives=prims, **vals)


def index_state(state: MPMState, b: int) -> MPMState:
    return tree_map_state(lambda t: t[b], state)


def stack_states(states) -> MPMState:
    s0 = states[0]
    prims = []
    for i in range(len(s0.primitives)):
        prims.append(P.PrimitiveState(*[torch.stack([getattr(s.primitives[i], k) for s in states])
                                        for k in P.PrimitiveState._fields]))
    vals = {k: (torch.stack([getattr(s, k) for s in states]) if getattr(s0, k) is not None else None)
            for k in s0._fields if k != "primitives"}
    return MPMState(primitives=prims, **vals)


def step_batch(sim: Simulator, state: MPMState, action: torch.Tensor) -> MPMState:
    """vmap(step) (mpm_simulator.py:61-63): independent per-env steps."""
    B = state.x.shape[0]
    return stack_states([sim.step(index_state(state, b), action[b]) for b in range(B)])


def reset_state(conf: MPMConf, x: torch.Tensor, prims, batch_size: int, dtype=torch.float32) -> MPMState:
    """reset, mpm_simulator.py:152-172 (PRNG key carried as zeros: pass-through only)."""
    n = x.shape[0]
    E, nu = conf.E, conf.nu
    mu_0, lambda_0 = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    s = MPMState(
        x=x.to(dtype), v=torch.zeros((n, 3), dtype=dtype), C=torch.zeros((n, 3, 3), dtype=dtype),
        F=torch.eye(3, dtype=dtype).reshape(1, 3, 3).repeat(n, 1, 1), J=torch.ones((n,), dtype=dtype),
        cur_step=torch.tensor(0, dtype=torch.int32), primitives=list(prims),
        key=torch.zeros(2, dtype=torch.int32),
        friction=torch.tensor([conf.ground_friction], dtype=dtype),
        mu=torch.tensor([mu_0], dtype=dtype), lamda=torch.tensor([lambda_0], dtype=dtype))
    return tree_map_state(lambda t: t[None].repeat((batch_size,) + (1,) * t.dim()), s)


def add_box(conf: MPMConf, size, init_pos, z_rotation_angle=0.0, density=1.0):
    """Lattice branch of add_box (material != 0), mpm_simulator.py:93-109.  Returns float32 x (n,3)."""
    import numpy as np
    size = np.asarray(size, dtype=np.float32)
    init_pos = np.asarray(init_pos, dtype=np.float32)
    ca, sa = np.float32(np.cos(z_rotation_angle)), np.float32(np.sin(z_rotation_angle))
    rot = np.array([[ca, -sa], [sa, ca]], dtype=np.float32)
    n_grid = int(conf.n_grid * density)
    center = np.array([0.5, 0.01, 0.5], dtype=np.float32)
    lower = (np.zeros(3, np.float32) * 2 - 1) * (np.float32(0.5) * size) + center
    upper = (np.ones(3, np.float32) * 2 - 1) * (np.float32(0.5) * size) + center
    a, b, c = np.indices((n_grid, n_grid, n_grid))
    gi = np.stack([a, b, c], axis=-1).astype(np.float32) * np.float32(1.0) / np.float32(n_grid)
    mask = np.all((gi <= upper) & (gi >= lower), axis=-1)
    x = gi[mask] - center
    x[:, [0, 2]] = x[:, [0, 2]] @ rot.T
    x = x + init_pos
    return torch.from_numpy(x.astype(np.float32))
