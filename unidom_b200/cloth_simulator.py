"""Host-side mirror of the reference's mass-spring cloth simulator interface, backed by sm_100a kernels.

Mirrors DaXBench/daxbench/core/engine/cloth_simulator.py: ClothState (:13-23),
ClothSimulator(conf, batch_size, collision_func, cloth_mask) (:26-70), .reset_jax (:339-364),
.step_jax (:107-180, one call = 50 substeps of one 8-vector sub-action), .get_x_grid (:366-368).
The env-level `collision_func` is the identity in every shipped task (cloth_env.py:239-243); a
non-identity one is rejected.  No CPU path.
"""
import ctypes as C
from typing import NamedTuple

import numpy as np
import torch

from . import _lib
from .mpm_simulator import _Workspace, _f32c, _ptr

LINKS = [[-1, 0], [1, 0], [0, -1], [0, 1], [-1, -1], [1, -1], [-1, 1], [1, 1]]  # cloth_simulator.py:48


class ClothState(NamedTuple):  # cloth_simulator.py:13-23
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


_LEAVES = ("x", "v", "primitive0", "primitive1", "action0", "action1", "stiffness", "mu")


class _ClothStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sim, stiff_float, action, *leaves):
        leaves = [_f32c(t) for t in leaves]
        action = _f32c(action)
        out = sim._call_fwd(leaves, action, stiff_float)
        ctx.sim, ctx.stiff_float = sim, stiff_float      # the dtype of THIS call's stiffness leaf, not the simulator's last
        ctx.save_for_backward(action, *leaves)
        return tuple(out)

    @staticmethod
    def backward(ctx, *gout):
        action, *leaves = ctx.saved_tensors
        gin, gaction = ctx.sim._call_bwd(leaves, action, list(gout), ctx.stiff_float)
        return (None, None, gaction, *gin)


class _ClothMultiStep(torch.autograd.Function):
    """One custom VJP for a whole scan over T sub-actions (cloth_env.py:211): forward = one launch, the adjoint
    recomputes from per-sub-action checkpoints inside the library."""

    @staticmethod
    def forward(ctx, sim, keep, stiff_float, actions, *leaves):
        leaves = [_f32c(t) for t in leaves]
        actions = _f32c(actions)
        out, ckpt = sim._call_multi_fwd(leaves, actions, keep, stiff_float)
        ctx.sim, ctx.ckpt, ctx.stiff_float = sim, ckpt, stiff_float
        ctx.save_for_backward(actions, *leaves)
        return tuple(out)

    @staticmethod
    def backward(ctx, *gout):
        actions, *leaves = ctx.saved_tensors
        gin, gactions = ctx.sim._call_multi_bwd(leaves, actions, ctx.ckpt, list(gout), ctx.stiff_float)
        return (None, None, None, gactions, *gin)


class ClothSimulator:
    """B200 drop-in for ClothSimulator (cloth_simulator.py:26-70)."""

    SUBSTEPS = 50  # cloth_simulator.py:176

    def __init__(self, conf, batch_size, collision_func=None, cloth_mask=None, device="cuda"):
        assert batch_size >= 1
        self._L = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("unidom_b200 needs a CUDA device (sm_100a); there is no CPU path")
        if collision_func is not None:
            probe = torch.zeros(1, 3)
            if collision_func(probe, probe, None, None) is not probe:
                raise NotImplementedError("only the identity env collision_func (cloth_env.py:239-243) is supported")
        self.conf = conf
        self.batch_size = batch_size
        self.device = torch.device(device)
        self.N = conf.N
        self.cell_size = 1.0 / conf.N
        self.cloth_mask = np.asarray(cloth_mask.cpu() if torch.is_tensor(cloth_mask) else cloth_mask)
        idx_i, idx_j = np.nonzero(self.cloth_mask)
        self.idx_i, self.idx_j = idx_i, idx_j
        self.n_nodes = len(idx_i)
        # topology tables (cloth_simulator.py:52-66)
        node_of = -np.ones((self.N, self.N), dtype=np.int64)
        node_of[idx_i, idx_j] = np.arange(self.n_nodes)
        grid_idx = np.stack([idx_i, idx_j], axis=-1)
        links = np.array(LINKS)
        j_ = np.clip(grid_idx[:, None, :] + links[None], 0, self.N - 1)
        i_ = np.repeat(grid_idx[:, None, :], 8, axis=1)
        ol = (np.float32(self.cell_size) * np.linalg.norm((j_ - i_).astype(np.float32), axis=-1)).astype(np.float32)
        nbr = node_of[j_[..., 0], j_[..., 1]]
        nbr = np.where((ol != 0) & (self.cloth_mask[j_[..., 0], j_[..., 1]] != 0), nbr, -1)
        self._nbr = torch.from_numpy(nbr.astype(np.int32)).to(self.device).contiguous()
        self._L0 = torch.from_numpy(np.clip(ol, 1e-12, np.inf).astype(np.float32)).to(self.device).contiguous()
        self._ws = _Workspace(self.device)
        x = np.zeros((self.N, self.N, 3))
        for i in range(self.N):
            x[i, :, 0] = i * self.cell_size
            x[i, :, 2] = (self.N - np.arange(self.N)) * self.cell_size
        self.x_grid = torch.from_numpy(x.astype(np.float32))

    def params(self, B, stiffness_is_float):
        c = self.conf
        p = _lib.ClothParams()
        p.num_envs, p.n_nodes, p.N, p.substeps = B, self.n_nodes, self.N, self.SUBSTEPS
        p.dt, p.gravity, p.damping = float(c.dt), float(c.gravity), float(c.damping)
        p.max_v, p.small_num, p.cell_size = float(c.max_v), float(c.small_num), float(self.cell_size)
        p.mask_sum = float(self.cloth_mask.sum())
        p.stiffness_is_float = int(stiffness_is_float)
        return p

    def reset_jax(self, stiffness=None, mu=None) -> ClothState:
        """cloth_simulator.py:339-364."""
        c = self.conf
        x = self.x_grid[self.idx_i, self.idx_j]
        B, dev = self.batch_size, self.device

        def rep(t):
            return t[None].repeat((B,) + (1,) * t.dim()).to(dev)
        stiff = torch.tensor(c.stiffness if stiffness is None else stiffness)
        return ClothState(
            x=rep(x), v=rep(torch.zeros_like(x)), primitive0=rep(torch.tensor([0.5, 0.5, 0.5, 0.01])),
            primitive1=rep(torch.tensor([1.0, 1.0, 1.0, 0.01])), action0=rep(torch.zeros(4)), action1=rep(torch.zeros(4)),
            key=rep(torch.zeros(2, dtype=torch.int32)), cur_step=rep(torch.tensor(0, dtype=torch.int32)),
            stiffness=rep(stiff), mu=rep(torch.tensor(float(c.mu if mu is None else mu))))

    def get_x_grid(self, x):
        g = self.x_grid.to(x.device)[None].repeat(x.shape[0], 1, 1, 1)
        g[:, self.idx_i, self.idx_j] = x
        return g

    def _pack(self, leaves):
        s = _lib.ClothState()
        for k, t in zip(_LEAVES, leaves):
            setattr(s, k, _ptr(t))
        return s

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call_fwd(self, leaves, action, stiff_float):
        p = self.params(leaves[0].shape[0], stiff_float)
        out = [torch.empty_like(t) for t in leaves]
        sin, sout = self._pack(leaves), self._pack(out)
        rc = self._L.ud_cloth_step_fwd(C.byref(p), C.byref(sin), _ptr(self._nbr), _ptr(self._L0), _ptr(action),
                                       C.byref(sout), C.c_void_p(0), 0, self._stream())
        _lib.check(rc, "ud_cloth_step_fwd")
        return out

    def _call_bwd(self, leaves, action, gout, stiff_float):
        p = self.params(leaves[0].shape[0], stiff_float)
        gout = [(_f32c(g) if g is not None else None) for g in gout]
        gin = [torch.zeros_like(t) for t in leaves]
        gaction = torch.zeros_like(action)
        ws, nbytes = self._ws.get(self._L.ud_cloth_workspace_bytes(C.byref(p)))
        rc = self._L.ud_cloth_step_bwd(C.byref(p), C.byref(self._pack(leaves)), _ptr(self._nbr), _ptr(self._L0),
                                       _ptr(action), C.byref(self._pack(gout)), C.byref(self._pack(gin)),
                                       _ptr(gaction), ws, nbytes, self._stream())
        _lib.check(rc, "ud_cloth_step_bwd")
        return gin, gaction

    def _call_multi_fwd(self, leaves, actions, keep, stiff_float):
        T = actions.shape[0]
        p = self.params(leaves[0].shape[0], stiff_float)
        out = [torch.empty_like(t) for t in leaves]
        ckpt, nbytes, cptr = None, 0, C.c_void_p(0)
        if keep:
            nbytes = self._L.ud_cloth_multi_ckpt_bytes(C.byref(p), T)
            ckpt = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            cptr = C.c_void_p(ckpt.data_ptr() + (-ckpt.data_ptr()) % 256)
        rc = self._L.ud_cloth_multi_step_fwd(C.byref(p), C.byref(self._pack(leaves)), _ptr(self._nbr), _ptr(self._L0),
                                             _ptr(actions), T, C.byref(self._pack(out)), cptr, nbytes, self._stream())
        _lib.check(rc, "ud_cloth_multi_step_fwd")
        return out, ckpt

    def _call_multi_bwd(self, leaves, actions, ckpt, gout, stiff_float):
        T = actions.shape[0]
        p = self.params(leaves[0].shape[0], stiff_float)
        gout = [(_f32c(g) if g is not None else None) for g in gout]
        gin = [torch.zeros_like(t) for t in leaves]
        gactions = torch.zeros_like(actions)
        ws, nbytes = self._ws.get(self._L.ud_cloth_multi_workspace_bytes(C.byref(p), T))
        cptr = C.c_void_p(ckpt.data_ptr() + (-ckpt.data_ptr()) % 256)
        rc = self._L.ud_cloth_multi_step_bwd(C.byref(p), C.byref(self._pack(leaves)), _ptr(self._nbr), _ptr(self._L0),
                                             _ptr(actions), T, cptr, C.byref(self._pack(gout)), C.byref(self._pack(gin)),
                                             _ptr(gactions), ws, nbytes, self._stream())
        _lib.check(rc, "ud_cloth_multi_step_bwd")
        return gin, gactions

    def scan_step_jax(self, state: ClothState, actions: torch.Tensor):
        """jax.lax.scan(self.step_jax, state, actions)[0] for actions [T,B,8] (cloth_env.py:211) as ONE fused call."""
        stiff_float = state.stiffness.is_floating_point()
        leaves = [getattr(state, k) for k in _LEAVES]
        leaves[6] = leaves[6].to(torch.float32)
        # grad mode is off inside Function.forward, so decide here whether the adjoint will need checkpoints
        keep = torch.is_grad_enabled() and (actions.requires_grad or any(t.requires_grad for t in leaves))
        out = _ClothMultiStep.apply(self, keep, stiff_float, actions, *leaves)
        vals = dict(zip(_LEAVES, out))
        if not stiff_float:
            vals["stiffness"] = state.stiffness
        return state._replace(**vals)

    def step_jax(self, state: ClothState, action: torch.Tensor):
        """vmap(jit(robot_step_wrapper)) (:68-70,107-180): returns (state, state)."""
        stiff_float = state.stiffness.is_floating_point()
        leaves = [getattr(state, k) for k in _LEAVES]
        leaves[6] = leaves[6].to(torch.float32)
        out = _ClothStep.apply(self, stiff_float, action, *leaves)
        vals = dict(zip(_LEAVES, out))
        if not stiff_float:
            vals["stiffness"] = state.stiffness
        new_state = state._replace(**vals)
        return new_state, new_state
