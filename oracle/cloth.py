"""Mass-spring cloth simulator -- torch-CPU restatement (test infrastructure).

Follows core/engine/cloth_simulator.py (line numbers cited per function).  Per environment; the
reference applies jax.vmap over axis 0 (:68-70), so the norms inside norm_grad's backward are
per-env.  Reverse mode = torch autograd of this forward with ``ClothNormGrad`` mirroring the
(effective, second) definition of norm_grad at :182-196.  The two custom_vjp recompute wrappers
(:107-145, :228-255) do not change the numbers except that cotangents of `key`/`cur_step` are dropped.
"""
from typing import NamedTuple

import numpy as np
import torch


class ClothState(NamedTuple):  # cloth_simulator.py:13-23 (same field order)
    x: torch.Tensor
    v: torch.Tensor
    primitive0: torch.Tensor
    primitive1: torch.Tensor
    action0: torch.Tensor
    action1: torch.Tensor
    key: torch.Tensor
    cur_step: torch.Tensor
    stiffness: torch.Tensor
    mu: torch.Tensor


class ClothConf:
    """fold_cloth3_env.py:18-37 defaults (shared by fold_cloth1/3, unfold_cloth1/3, fold_cloth1_para)."""
    N = 80
    gravity = 0.5
    stiffness = 900
    damping = 2
    dt = 2e-3
    max_v = 2.0
    small_num = 1e-8
    mu = 0.5
    seed = 1
    mem_saving_level = 2

    @property
    def cell_size(self):
        return 1.0 / self.N

    @property
    def size(self):
        return int(self.N / 5.0)


def fold_cloth_mask(conf):
    """fold_cloth3_env.py:51-56 (identical in fold_cloth1 / fold_cloth1_para / unfold_cloth*)."""
    N, size = conf.N, conf.size
    m = np.zeros((N, N), dtype=np.float32)
    m[size * 2:size * 3, size * 2:size * 4] = 1
    return m


LINKS = [[-1, 0], [1, 0], [0, -1], [0, 1], [-1, -1], [1, -1], [-1, 1], [1, 1]]  # :48


def jclip(x, lo=None, hi=None):
    """jnp.clip = minimum(hi, maximum(lo, x)) (jax/_src/numpy/lax_numpy.py): at a tie with a bound the
    cotangent is split evenly (lax.max/min JVP), where torch.clamp would pass all of it.  Cloth nodes rest
    at y == 0 exactly after reset and the idle gripper sits at 1.0 exactly, so ties do occur."""
    if lo is not None:
        x = torch.maximum(x, torch.as_tensor(lo, dtype=x.dtype))
    if hi is not None:
        x = torch.minimum(x, torch.as_tensor(hi, dtype=x.dtype))
    return x


class ClothNormGrad(torch.autograd.Function):
    """norm_grad (cloth_simulator.py:182-196): identity forward; backward g/|g|, nan_to_num, /mask.sum()."""

    @staticmethod
    def forward(ctx, x, mask_sum):
        ctx.mask_sum = mask_sum
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        g = g / torch.linalg.norm(g)
        g = torch.nan_to_num(g)
        g = g / ctx.mask_sum
        return g, None


class ClothSim:
    """Per-environment functional core of ClothSimulator (:26-70)."""

    def __init__(self, conf, cloth_mask, dtype=torch.float32):
        self.conf = conf
        self.dtype = dtype
        self.N = conf.N
        self.cell_size = 1.0 / conf.N
        cloth_mask = np.asarray(cloth_mask)
        self.cloth_mask = torch.from_numpy(cloth_mask.astype(np.float32)).to(dtype)
        idx_i, idx_j = np.nonzero(cloth_mask)
        self.idx_i, self.idx_j = torch.from_numpy(idx_i), torch.from_numpy(idx_j)
        grid_idx = np.stack([idx_i, idx_j], axis=-1)
        links = np.array(LINKS)
        j_ = np.clip(grid_idx[:, None, :] + links[None], 0, self.N - 1)                 # :56-58
        i_ = np.repeat(grid_idx[:, None, :], 8, axis=1)
        # :61 cell_size (python float) * norm (f32) -> f32
        ol = (np.float32(self.cell_size) * np.linalg.norm((j_ - i_).astype(np.float32), axis=-1)[..., None]).astype(np.float32)
        self.ori_len_is_not_0 = torch.from_numpy((ol != 0).astype(np.float32)).to(dtype)  # (P,8,1)
        self.original_length = torch.from_numpy(np.clip(ol, 1e-12, np.inf)).to(dtype)
        self.j_x = torch.from_numpy(j_.reshape(-1, 2)[:, 0].copy())
        self.j_y = torch.from_numpy(j_.reshape(-1, 2)[:, 1].copy())
        self.i_x = torch.from_numpy(i_.reshape(-1, 2)[:, 0].copy())
        self.i_y = torch.from_numpy(i_.reshape(-1, 2)[:, 1].copy())
        self.mask_sum = float(cloth_mask.sum())
        self.n_nodes = len(idx_i)

    def norm_grad(self, t):
        return ClothNormGrad.apply(t, self.mask_sum)

    def primitive_collision(self, x, v, action, ps):
        """:198-226."""
        pos, radius = ps[:3], ps[3]
        d_v = action[:3].reshape(1, 3)
        suction = action[-1]
        dist = torch.linalg.norm(x - pos.reshape(1, 3), dim=-1)
        mask = (dist <= radius)[..., None].expand(-1, 3)
        v_ = torch.where(mask, suction * v, v)
        x_ = torch.where(mask, x + d_v * (1 - suction), x)
        return self.norm_grad(x_), self.norm_grad(v_)

    def step(self, state: ClothState) -> ClothState:
        """:257-337."""
        c = self.conf
        dtype = state.x.dtype
        x, v = state.x, state.v
        v = v - torch.tensor([0.0, c.gravity * c.dt, 0.0], dtype=dtype)
        x_grid = torch.zeros((self.N, self.N, 3), dtype=dtype).index_put((self.idx_i, self.idx_j), x)
        rel = x_grid[self.j_x, self.j_y] - x_grid[self.i_x, self.i_y]
        cur = jclip((rel ** 2).sum(-1), 1e-12) ** 0.5
        cur = cur.reshape(-1, 8, 1)
        force = state.stiffness * rel.reshape(-1, 8, 3) / cur * (cur - self.original_length) / self.original_length
        force = force * self.ori_len_is_not_0
        force = force * self.cloth_mask[self.j_x, self.j_y].reshape(-1, 8, 1)
        force = force.sum(1)
        fx, fy, fz = force[:, 0], force[:, 1] - c.gravity, force[:, 2]
        # friction (:281-306)
        friction_mask = x[:, 1] <= c.small_num
        muF = state.mu * jclip(fy, None, 0.0) * -1
        xV, yV = v[:, 0], v[:, 2]
        sV = torch.sqrt(xV ** 2 + yV ** 2 + c.small_num)
        dyn = (friction_mask & (sV > c.small_num)).to(dtype)
        fx = fx - dyn * muF * xV / sV
        fz = fz - dyn * muF * yV / sV
        static = friction_mask & (sV <= c.small_num)
        xF, yF = fx, fz
        sF = torch.sqrt(xF ** 2 + yF ** 2 + c.small_num)
        zero = (static & (muF > sF)).to(dtype)
        fx = 0 + (1.0 - zero) * fx
        fz = 0 + (1.0 - zero) * fz
        nonzero = (static & (muF <= sF)).to(dtype)
        R = 1.0 - muF / sF
        fx = (R * xF) * nonzero + fx * (1.0 - nonzero)
        fz = (R * yF) * nonzero + fz * (1.0 - nonzero)
        force = torch.stack([fx, fy, fz], dim=-1)
        v = v + force * c.dt
        v = v * float(np.exp(np.float32(-c.damping * c.dt)))
        # collision_func is the identity for all shipped tasks (cloth_env.py:239-243)
        x, v = self.primitive_collision(x, v, state.action0, state.primitive0)
        x, v = self.primitive_collision(x, v, state.action1, state.primitive1)
        ps0 = jclip(torch.cat([state.primitive0[:3] + state.action0[:3], state.primitive0[3:]]), 0, 1)
        ps1 = jclip(torch.cat([state.primitive1[:3] + state.action1[:3], state.primitive1[3:]]), 0, 1)
        x = jclip(x, 0, 1)
        v = jclip(v, -c.max_v, c.max_v)
        x = x + c.dt * v
        x, v = self.norm_grad(x), self.norm_grad(v)
        ps0, ps1 = self.norm_grad(ps0), self.norm_grad(ps1)
        return state._replace(x=x, v=v, primitive0=ps0, primitive1=ps1)

    def robot_step(self, state: ClothState, action: torch.Tensor, substeps: int = 50) -> ClothState:
        """:163-180 (50 substeps; the PRNG key split consumes no randomness and is carried as is)."""
        action0 = torch.cat([jclip(action[:3], -2, 2) / 50.0, action[3:4]])
        action1 = torch.cat([jclip(action[4:7], -2, 2) / 50.0, action[7:8]])
        state = state._replace(action0=action0, action1=action1)
        for _ in range(substeps):
            state = self.step(state)
        return state

    def reset(self, batch_size, stiffness=None, mu=None):
        """:339-364 (lattice x, zero v, grippers at (0.5,0.5,0.5) / (1,1,1), radius 0.01)."""
        c = self.conf
        N = self.N
        xg = np.zeros((N, N, 3))
        for i in range(N):
            for j in range(N):
                xg[i, j] = np.array([i * self.cell_size, 0, (N - j) * self.cell_size])
        x = torch.from_numpy(xg.astype(np.float32))[self.idx_i, self.idx_j].to(self.dtype)
        st = ClothState(
            x=x, v=torch.zeros_like(x),
            primitive0=torch.tensor([0.5, 0.5, 0.5, 0.01], dtype=self.dtype),
            primitive1=torch.tensor([1.0, 1.0, 1.0, 0.01], dtype=self.dtype),
            action0=torch.zeros(4, dtype=self.dtype), action1=torch.zeros(4, dtype=self.dtype),
            key=torch.zeros(2, dtype=torch.int32), cur_step=torch.tensor(0, dtype=torch.int32),
            stiffness=torch.tensor(float(c.stiffness if stiffness is None else stiffness), dtype=self.dtype),
            mu=torch.tensor(float(c.mu if mu is None else mu), dtype=self.dtype))
        return ClothState(*[t[None].repeat((batch_size,) + (1,) * t.dim()) for t in st])


def index_state(state: ClothState, b: int) -> ClothState:
    return ClothState(*[t[b] for t in state])


def stack_states(states) -> ClothState:
    return ClothState(*[torch.stack([getattr(s, k) for s in states]) for k in ClothState._fields])


def step_batch(sim: ClothSim, state: ClothState, action: torch.Tensor, substeps: int = 50) -> ClothState:
    """vmap(robot_step_wrapper) (:68-70)."""
    B = state.x.shape[0]
    return stack_states([sim.robot_step(index_state(state, b), action[b], substeps) for b in range(B)])


def get_pnp_actions(actions: torch.Tensor, state: ClothState) -> torch.Tensor:
    """cloth_env.py:136-173, batched: actions (B,6) -> sub-actions (40, B, 8)."""
    outs = []
    for b in range(actions.shape[0]):
        a = actions[b]
        pick, place = a[:3].clone(), a[3:].clone()
        pick = torch.cat([pick[:1], torch.zeros(1, dtype=a.dtype), pick[2:]])
        place = torch.cat([place[:1], torch.zeros(1, dtype=a.dtype), place[2:]])
        act_down = torch.cat([(pick - state.primitive0[b, :3]) / 3, torch.ones(1, dtype=a.dtype)])[None].repeat(3, 1)
        act_up = torch.tensor([0, 0.06, 0, 0], dtype=a.dtype)[None].repeat(10, 1)
        act_up = torch.cat([act_up[:, :3] / 10, act_up[:, 3:]], dim=1)
        mv = place - pick
        mv = torch.cat([mv[:1], torch.zeros(1, dtype=a.dtype), mv[2:], torch.zeros(1, dtype=a.dtype)])
        act_move = torch.cat([mv[None, :3].repeat(20, 1) / 20, mv[None, 3:].repeat(20, 1)], dim=1)
        act_release = torch.tensor([0, 0, 0, 1], dtype=a.dtype)[None].repeat(7, 1)
        sub = torch.cat([act_down, act_up, act_move, act_release], dim=0)
        outs.append(torch.cat([sub, torch.zeros_like(sub)], dim=1))
    return torch.stack(outs, dim=1)
