"""Host-side mirror of the reference's MLS-MPM simulator interface, backed by the sm_100a kernels.

Mirrors DaXBench/daxbench/core/engine/mpm_simulator.py:
    MPMState (:13-24), SimpleMPMSimulator(conf, batch_size, use_position_control) (:27-63),
    .add_box (:65-125), .add_box_from_points (:127-145), .reset_jax (:152-172), .step_jax (:413-429)
and primitives.py: PrimitiveState (:9-23), create_primitive (:31-60).

Device memory, streams and autograd plumbing are torch; all arithmetic of the step runs in
libunidom_b200.so through the C ABI (include/unidom_b200.h).  There is no CPU path: construction
fails if the library is missing or CUDA is unavailable.
"""
import ctypes as C
import math
from typing import List, NamedTuple

import numpy as np
import torch

from . import _lib


class PrimitiveState(NamedTuple):  # primitives.py:9-23
    size: torch.Tensor
    dim: torch.Tensor
    friction: torch.Tensor
    softness: torch.Tensor
    color: torch.Tensor
    position: torch.Tensor
    rotation: torch.Tensor
    v: torch.Tensor
    w: torch.Tensor
    xyz_limit: torch.Tensor
    action_buffer: torch.Tensor
    action_scale: torch.Tensor
    min_dist: torch.Tensor
    dist_norm: torch.Tensor


class MPMState(NamedTuple):  # mpm_simulator.py:13-24
    x: torch.Tensor = None
    v: torch.Tensor = None
    C: torch.Tensor = None
    F: torch.Tensor = None
    J: torch.Tensor = None
    cur_step: torch.Tensor = None
    primitives: List[PrimitiveState] = []
    key: torch.Tensor = None
    friction: torch.Tensor = None
    mu: torch.Tensor = None
    lamda: torch.Tensor = None


def create_primitive(conf, friction, softness, color, size, init_pos):
    """primitives.py:31-60 (unbatched, host tensors)."""
    steps = conf.steps
    position = torch.zeros((steps, 3))
    position[0] = torch.as_tensor(init_pos, dtype=torch.float32)
    return PrimitiveState(
        size=torch.as_tensor(np.asarray(size), dtype=torch.float32),
        dim=torch.tensor([3], dtype=torch.int32),
        friction=torch.tensor(float(friction)),
        softness=torch.tensor(float(softness)),
        color=torch.as_tensor(np.asarray(color), dtype=torch.float32),
        position=position,
        rotation=torch.tensor([[1.0, 0.0, 0.0, 0.0]]).repeat(steps, 1),
        v=torch.zeros((steps, 3)), w=torch.zeros((steps, 3)),
        xyz_limit=torch.tensor([[0.0, 1.0]] * 3),
        action_buffer=torch.zeros((6,)), action_scale=torch.ones((6,)),
        min_dist=torch.tensor(0, dtype=torch.int32), dist_norm=torch.tensor(0, dtype=torch.int32))


# differentiable leaves, in the order the autograd Function flattens them
_STATE_LEAVES = ("x", "v", "C", "F", "J", "friction", "mu", "lamda")
_PRIM_LEAVES = ("size", "friction", "position", "rotation", "v", "w", "action_buffer", "action_scale")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _aligned(buf):
    return C.c_void_p(buf.data_ptr() + (-buf.data_ptr()) % 256)


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class _Workspace:
    """Caller-owned scratch for the C ABI (never allocated inside the library)."""

    def __init__(self, device):
        self.device = device
        self.buf = None
        self.pins = 0       # live CUDA graphs that captured the pointer of `buf` (graphs.GraphedMPMScan)

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            if self.pins:
                raise RuntimeError(
                    f"this simulator's workspace ({0 if self.buf is None else self.buf.numel()} B) is captured by a live "
                    f"CUDA graph and cannot grow to {nbytes} B: use a second simulator for the larger batch, or drop the "
                    "graph (GraphedMPMScan.close()) first")
            self.buf = None
            self.buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        base = self.buf.data_ptr()
        off = (-base) % 256
        return C.c_void_p(base + off), nbytes


def flatten_state(state: MPMState, n_prim: int):
    leaves = [getattr(state, k) for k in _STATE_LEAVES]
    for q in range(n_prim):
        leaves += [getattr(state.primitives[q], k) for k in _PRIM_LEAVES]
    return leaves


class _Tape:
    """One step's substep residuals.  Buffers are recycled through the simulator's own free list: a 6 GB request per
    differentiated step that goes through torch's caching allocator gets its cached block split by the small
    allocations in between and ends in cudaMalloc/cudaFree churn (measured: 11 -> 16-29 ms per step, erratic)."""

    def __init__(self, sim, nbytes):
        self.sim, self.nbytes = sim, nbytes
        pool = sim._tape_pool
        i = next((i for i, b in enumerate(pool) if b.numel() == nbytes + 256), None)
        self.buf = pool.pop(i) if i is not None else torch.empty(nbytes + 256, dtype=torch.uint8, device=sim.device)

    def release(self):
        """Called after the step's backward: the buffer goes back to the free list (a second backward through the same
        step needs adjoint="recompute", like retain_graph needs saved tensors)."""
        if self.buf is not None:
            if len(self.sim._tape_pool) < self.sim.tape_pool_size:
                self.sim._tape_pool.append(self.buf)
            self.buf = None

    def __del__(self):      # forward that never saw a backward
        try:
            self.release()
        except Exception:
            pass


class _MpmStep(torch.autograd.Function):
    """One custom VJP per (env batch, sub-action): fwd = S substeps; bwd = reverse from the tape the forward kept,
    or recompute + reverse when the step ran without a tape
    (replaces substep_wrapper / norm_grad* custom_vjps, mpm_simulator.py:332-411)."""

    @staticmethod
    def forward(ctx, sim, softness_list, want_tape, action, *leaves):
        leaves = [_f32c(t) for t in leaves]
        action = _f32c(action)
        out_leaves, ctx.tape = sim._call_fwd(leaves, softness_list, action, want_tape)
        ctx.sim = sim
        ctx.softness_list = softness_list
        ctx.save_for_backward(action, *leaves)
        return tuple(out_leaves)

    @staticmethod
    def backward(ctx, *gout):
        action, *leaves = ctx.saved_tensors
        tape = ctx.tape
        if tape is not None and tape.buf is None:
            raise RuntimeError("this step's tape was released by an earlier backward; differentiate it once, or build "
                               "the simulator with adjoint='recompute'")
        gin, gaction = ctx.sim._call_bwd(leaves, ctx.softness_list, action, list(gout), tape)
        if tape is not None:
            tape.release()
        return (None, None, None, gaction, *gin)


class SimpleMPMSimulator:
    """B200 drop-in for SimpleMPMSimulator (mpm_simulator.py:27-63)."""

    def __init__(self, conf, batch_size, use_position_control=False, device="cuda", sdf_kind=None,
                 p2g_mode=_lib.UD_P2G_ATOMIC, adjoint="recompute", tape_budget_bytes=None, ckpt_window=None,
                 env_groups=1):
        """adjoint: "recompute" (default) keeps only the step input and re-runs the S substeps in the backward;
        "tape" keeps every substep's residuals of a differentiated step in HBM until its backward (what jax.grad of
        the reference's lax.scan does: S times the memory, 0.71x the time); "auto" tapes while this process's allocated device memory plus
        the new tape stays within `tape_budget_bytes` (default: 70 % of the device's memory) and recomputes beyond.
        env_groups: the envs of a call are independent (vmap over axis 0), so a call may run them as `env_groups`
        sub-batches on that many CUDA streams: the latency-bound grid / binning kernels of one group overlap the
        particle kernels of the others.  Results are those of env_groups=1 (every env is computed by the same kernels).
        ckpt_window: K-spaced substep checkpoints inside the recompute adjoint (ud_mpm_step_bwd_windowed): the backward
        keeps the state entering every K-th substep and recomputes K substeps at a time, so its workspace holds K instead
        of conf.steps substeps (long steps: whip_rope's 70 substeps) for one extra forward sweep; None = conf.steps."""
        self._L = _lib.lib()  # raises when the extension is missing
        if not torch.cuda.is_available():
            raise RuntimeError("unidom_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.conf = conf
        self.batch_size = batch_size
        self.use_position_control = use_position_control
        self.device = torch.device(device)
        self.n_particles = 0
        self.material = None  # host int32 [n]   (mpm_simulator.py:57,117)
        self._liq_cache = (None, False)
        self.h = None         # host float32 [n] (mpm_simulator.py:58,118)
        self.key_global = None
        if sdf_kind is None:
            sdf_kind = getattr(conf, "sdf_kind", _lib.UD_SDF_BOX)
        self.sdf_kind = int(sdf_kind)
        self.p2g_mode = int(p2g_mode)
        self.env_groups = max(1, int(env_groups))
        self._ws_fwd_g = [_Workspace(self.device) for _ in range(self.env_groups)]
        self._ws_bwd_g = [_Workspace(self.device) for _ in range(self.env_groups)]
        self._ws_fwd, self._ws_bwd = self._ws_fwd_g[0], self._ws_bwd_g[0]
        self._group_streams = None
        if adjoint not in ("auto", "tape", "recompute"):
            raise ValueError(f"adjoint={adjoint!r}: expected 'auto', 'tape' or 'recompute'")
        self.adjoint = adjoint
        if ckpt_window is not None and int(ckpt_window) < 1:
            raise ValueError(f"ckpt_window={ckpt_window!r}: expected a positive number of substeps")
        self.ckpt_window = None if ckpt_window is None else int(ckpt_window)
        if tape_budget_bytes is None:
            tape_budget_bytes = torch.cuda.get_device_properties(self.device).total_memory * 7 // 10
        self.tape_budget_bytes = int(tape_budget_bytes)
        self.last_adjoint = None        # "tape" / "recompute": what the last differentiated forward chose
        self._tape_pool = []            # free tape buffers (see _Tape)
        self.tape_pool_size = 2
        self._rng = np.random.RandomState(getattr(conf, "seed", 0))

    # ------------------------------------------------------------- scene construction
    def add_box(self, conf, state, size, init_pos, hardness=1, z_rotation_angle=0, material=0, density=1):
        """mpm_simulator.py:65-125.  The liquid branch draws jax.random.uniform(conf.key, (n_points, 3)) through the
        NumPy threefry of unidom_b200.jaxrng when the conf carries a `key` (a PRNGKey or an int seed, like the
        reference's confs: `key = random.PRNGKey(0)`); without one, a seeded NumPy stream."""
        assert density >= 1
        size = np.asarray(size, dtype=np.float32)
        init_pos = np.asarray(init_pos, dtype=np.float32)
        ca, sa = np.float32(math.cos(z_rotation_angle)), np.float32(math.sin(z_rotation_angle))
        rot = np.array([[ca, -sa], [sa, ca]], dtype=np.float32)
        if material == 0:
            n_points = int(np.prod(size.astype(np.float64)) * conf.n_grid ** 3 * density)
            key = getattr(conf, "key", None)
            if key is not None:
                from . import jaxrng
                key = jaxrng.PRNGKey(key) if np.isscalar(key) else np.asarray(key, np.uint32)
                u = jaxrng.uniform(key, (n_points, 3))                    # mpm_simulator.py:89, same stream
            else:
                u = self._rng.uniform(size=(n_points, 3)).astype(np.float32)
            x_ = (u * np.float32(2) - np.float32(1)) * (np.float32(0.5) * size)
            x_[:, [0, 2]] = x_[:, [0, 2]] @ rot.T
            x_ = x_ + init_pos
        else:
            n_grid = int(conf.n_grid * density)
            center = np.array([0.5, 0.01, 0.5], dtype=np.float32)
            lower = -(np.float32(0.5) * size) + center
            upper = (np.float32(0.5) * size) + center
            # the reference masks a full (n_grid^3, 3) index lattice; the mask is separable per axis, so only the
            # 1-D candidates are formed here (same float32 values, same row-major order of the surviving points)
            ax = np.arange(n_grid).astype(np.float32) * np.float32(1.0) / np.float32(n_grid)
            keep = [ax[(ax <= upper[d]) & (ax >= lower[d])] for d in range(3)]
            a, b, c = np.meshgrid(keep[0], keep[1], keep[2], indexing="ij")
            x_ = np.stack([a, b, c], axis=-1).reshape(-1, 3) - center
            x_[:, [0, 2]] = x_[:, [0, 2]] @ rot.T
            x_ = x_ + init_pos
        return self.add_box_from_points(conf, state, torch.from_numpy(x_.astype(np.float32)), hardness, material)

    def add_box_from_points(self, conf, state, points, hardness=1, material=0):
        """mpm_simulator.py:127-145."""
        x_ = torch.as_tensor(points, dtype=torch.float32)
        n_points = x_.shape[0]
        material_ = torch.full((n_points,), int(material), dtype=torch.int32)
        h_ = torch.full((n_points,), float(hardness), dtype=torch.float32)
        if state is None:
            self.material, self.h = material_, h_
            prims = []
        else:
            x_ = torch.cat([state.x.cpu(), x_], dim=0)
            self.material = torch.cat([self.material, material_])
            self.h = torch.cat([self.h, h_])
            prims = list(state.primitives)
        return MPMState(x=x_, primitives=prims)

    def reset_jax(self, state: MPMState) -> MPMState:
        """mpm_simulator.py:152-172: broadcast to the batch and move to the device."""
        conf = self.conf
        self.n_particles = n = state.x.shape[0]
        E, nu = conf.E, conf.nu
        mu_0, lambda_0 = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
        B, dev = self.batch_size, self.device

        def rep(t):
            t = torch.as_tensor(t)
            return t[None].repeat((B,) + (1,) * t.dim()).to(dev)

        prims = [PrimitiveState(*[rep(t) for t in p]) for p in state.primitives]
        self._material_dev = self.material.to(dev).contiguous()
        self._h_dev = self.h.to(dev).contiguous()
        return MPMState(
            x=rep(state.x.to(torch.float32)), v=rep(torch.zeros((n, 3))), C=rep(torch.zeros((n, 3, 3))),
            F=rep(torch.eye(3).reshape(1, 3, 3).repeat(n, 1, 1)), J=rep(torch.ones((n,))),
            cur_step=rep(torch.tensor(0, dtype=torch.int32)), primitives=prims,
            key=self._batch_keys(B).to(dev),
            friction=rep(torch.tensor([float(conf.ground_friction)])),
            mu=rep(torch.tensor([mu_0], dtype=torch.float32)),
            lamda=rep(torch.tensor([lambda_0], dtype=torch.float32)))

    def _batch_keys(self, B):
        """mpm_simulator.py:169: state.key = jax.random.split(key_global, batch_size) (uint32 bit patterns kept in the
        int32 leaf); zeros when no key_global was set (the reference fails there)."""
        if self.key_global is None:
            return torch.zeros((B, 2), dtype=torch.int32)
        from . import jaxrng
        kg = jaxrng.PRNGKey(self.key_global) if np.isscalar(self.key_global) else np.asarray(self.key_global, np.uint32)
        return torch.from_numpy(jaxrng.split(kg, B).view(np.int32).copy())

    # ------------------------------------------------------------------------ the step
    def _has_liquid(self):
        """Material-0 particles in the scene -> UD_P2G_LIQUID_FAST (kernels with the SVD-free liquid path).  Only a
        speed hint: without the flag liquid particles take the general path, with it other scenes pay a branch."""
        m = self.material
        if m is None:
            return False
        if self._liq_cache[0] != id(m):
            self._liq_cache = (id(m), bool((torch.as_tensor(m).cpu() == 0).any()))
        return self._liq_cache[1]

    def params(self, B=None, n=None):
        conf = self.conf
        p = _lib.MpmParams()
        p.num_envs = self.batch_size if B is None else B
        p.n_particles = self.n_particles if n is None else n
        p.steps = int(conf.steps)
        p.res = (C.c_int32 * 3)(*[int(r) for r in conf.res])
        p.n_grid = int(conf.n_grid)
        p.dt, p.dx, p.inv_dx = float(conf.dt), float(conf.dx), float(conf.inv_dx)
        p.p_mass, p.p_vol = float(conf.p_mass), float(conf.p_vol)
        p.gravity = (C.c_double * 3)(*[float(g) for g in np.asarray(conf.gravity, dtype=np.float64)])
        p.n_primitive = int(conf.n_primitive)
        p.sdf_kind = self.sdf_kind
        p.use_position_control = int(bool(self.use_position_control))
        p.p2g_mode = self.p2g_mode | (_lib.UD_P2G_LIQUID_FAST if self._has_liquid() else 0)
        return p

    def _pack(self, leaves, softness_list):
        """Flattened leaves -> ud_mpm_state (keeps the tensors alive through `leaves`)."""
        s = _lib.MpmState()
        for i, k in enumerate(_STATE_LEAVES):
            setattr(s, k, _ptr(leaves[i]))
        n0 = len(_STATE_LEAVES)
        nprim = (len(leaves) - n0) // len(_PRIM_LEAVES)
        for q in range(nprim):
            pl = leaves[n0 + q * len(_PRIM_LEAVES): n0 + (q + 1) * len(_PRIM_LEAVES)]
            for j, k in enumerate(_PRIM_LEAVES):
                setattr(s.prim[q], k, _ptr(pl[j]))
            s.prim[q].softness = _ptr(softness_list[q]) if softness_list is not None else C.c_void_p(0)
        return s

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _groups(self, B):
        """[(first env, envs, stream, group index)] of a call over B envs; one group on the current stream unless
        env_groups > 1 divides the batch."""
        G = self.env_groups
        if G <= 1 or B % G or B < G:
            return None
        if self._group_streams is None:
            self._group_streams = [torch.cuda.Stream(self.device) for _ in range(G)]
        per = B // G
        return [(g * per, per, self._group_streams[g], g) for g in range(G)]

    def _call_fwd(self, leaves, softness_list, action, want_tape=False):
        p = self.params(B=leaves[0].shape[0], n=leaves[0].shape[1])
        out = [torch.empty_like(t) for t in leaves]
        soft_out = [torch.empty_like(t) for t in softness_list]
        groups = self._groups(leaves[0].shape[0])
        if groups is not None and not (want_tape and self.adjoint != "recompute"):
            if want_tape:
                self.last_adjoint = "recompute"
            cur = torch.cuda.current_stream(self.device)
            for e0, ne, sg, g in groups:
                sg.wait_stream(cur)
                with torch.cuda.stream(sg):
                    pg = self.params(B=ne, n=leaves[0].shape[1])
                    sl = lambda ts: [t[e0:e0 + ne] for t in ts]      # noqa: E731 -- batch is the leading axis of every leaf
                    sin, sout = self._pack(sl(leaves), sl(softness_list)), self._pack(sl(out), sl(soft_out))
                    ws, nbytes = self._ws_fwd_g[g].get(self._L.ud_mpm_fwd_workspace_bytes(C.byref(pg)))
                    rc = self._L.ud_mpm_step_fwd(C.byref(pg), C.byref(sin), _ptr(self._material_dev), _ptr(self._h_dev),
                                                 _ptr(action[e0:e0 + ne]), C.byref(sout), ws, nbytes,
                                                 C.c_void_p(sg.cuda_stream))
                    _lib.check(rc, "ud_mpm_step_fwd")
            for _, _, sg, _ in groups:
                cur.wait_stream(sg)
            return out, None
        sin, sout = self._pack(leaves, softness_list), self._pack(out, soft_out)
        if want_tape and self.adjoint != "recompute":
            need = self._L.ud_mpm_tape_bytes(C.byref(p))
            pooled = any(b.numel() == need + 256 for b in self._tape_pool)      # already counted in memory_allocated
            if (self.adjoint == "tape" or pooled
                    or torch.cuda.memory_allocated(self.device) + need <= self.tape_budget_bytes):
                tape = _Tape(self, need)
                rc = self._L.ud_mpm_step_fwd_taped(C.byref(p), C.byref(sin), _ptr(self._material_dev),
                                                   _ptr(self._h_dev), _ptr(action), C.byref(sout), _aligned(tape.buf),
                                                   need, self._stream())
                _lib.check(rc, "ud_mpm_step_fwd_taped")
                self.last_adjoint = "tape"
                return out, tape
        if want_tape:
            self.last_adjoint = "recompute"
        ws, nbytes = self._ws_fwd.get(self._L.ud_mpm_fwd_workspace_bytes(C.byref(p)))
        rc = self._L.ud_mpm_step_fwd(C.byref(p), C.byref(sin), _ptr(self._material_dev), _ptr(self._h_dev),
                                     _ptr(action), C.byref(sout), ws, nbytes, self._stream())
        _lib.check(rc, "ud_mpm_step_fwd")
        return out, None

    def _call_bwd(self, leaves, softness_list, action, gout, tape=None):
        p = self.params(B=leaves[0].shape[0], n=leaves[0].shape[1])
        gout = [(_f32c(g) if g is not None else None) for g in gout]
        gin = [torch.zeros_like(t) for t in leaves]
        gaction = torch.zeros_like(action)
        sin = self._pack(leaves, softness_list)
        sgo, sgi = self._pack(gout, None), self._pack(gin, None)
        if tape is not None:
            rc = self._L.ud_mpm_step_bwd_taped(C.byref(p), C.byref(sin), _ptr(action), C.byref(sgo), C.byref(sgi),
                                               _ptr(gaction), _aligned(tape.buf), tape.nbytes, self._stream())
            _lib.check(rc, "ud_mpm_step_bwd_taped")
            return gin, gaction
        groups = self._groups(leaves[0].shape[0])
        if groups is not None:
            cur = torch.cuda.current_stream(self.device)
            K = self.ckpt_window if (self.ckpt_window is not None and self.ckpt_window < int(self.conf.steps)) else int(self.conf.steps)
            for e0, ne, sg, g in groups:
                sg.wait_stream(cur)
                with torch.cuda.stream(sg):
                    pg = self.params(B=ne, n=leaves[0].shape[1])
                    sl = lambda ts: [(t[e0:e0 + ne] if t is not None else None) for t in ts]      # noqa: E731
                    s_in, s_go, s_gi = self._pack(sl(leaves), sl(softness_list)), self._pack(sl(gout), None), self._pack(sl(gin), None)
                    ws, nbytes = self._ws_bwd_g[g].get(self._L.ud_mpm_bwd_windowed_workspace_bytes(C.byref(pg), K))
                    rc = self._L.ud_mpm_step_bwd_windowed(C.byref(pg), C.byref(s_in), _ptr(self._material_dev),
                                                          _ptr(self._h_dev), _ptr(action[e0:e0 + ne]), C.byref(s_go),
                                                          C.byref(s_gi), _ptr(gaction[e0:e0 + ne]), K, ws, nbytes,
                                                          C.c_void_p(sg.cuda_stream))
                    _lib.check(rc, "ud_mpm_step_bwd_windowed")
            for _, _, sg, _ in groups:
                cur.wait_stream(sg)
            return gin, gaction
        if self.ckpt_window is not None and self.ckpt_window < int(self.conf.steps):
            K = self.ckpt_window
            ws, nbytes = self._ws_bwd.get(self._L.ud_mpm_bwd_windowed_workspace_bytes(C.byref(p), K))
            rc = self._L.ud_mpm_step_bwd_windowed(C.byref(p), C.byref(sin), _ptr(self._material_dev), _ptr(self._h_dev),
                                                  _ptr(action), C.byref(sgo), C.byref(sgi), _ptr(gaction), K, ws, nbytes,
                                                  self._stream())
            _lib.check(rc, "ud_mpm_step_bwd_windowed")
            return gin, gaction
        ws, nbytes = self._ws_bwd.get(self._L.ud_mpm_bwd_workspace_bytes(C.byref(p)))
        rc = self._L.ud_mpm_step_bwd(C.byref(p), C.byref(sin), _ptr(self._material_dev), _ptr(self._h_dev),
                                     _ptr(action), C.byref(sgo), C.byref(sgi), _ptr(gaction), ws, nbytes,
                                     self._stream())
        _lib.check(rc, "ud_mpm_step_bwd")
        return gin, gaction

    def step_jax(self, state: MPMState, action: torch.Tensor):
        """vmap(jit(step)) (mpm_simulator.py:61-63,413-429): returns (state, state)."""
        n_prim = int(self.conf.n_primitive)
        leaves = flatten_state(state, n_prim)
        softness = [_f32c(state.primitives[q].softness) for q in range(n_prim)]
        # grad mode is off inside Function.forward: decide here whether a backward can follow
        want_tape = torch.is_grad_enabled() and (action.requires_grad or any(t.requires_grad for t in leaves))
        out = _MpmStep.apply(self, softness, want_tape, action, *leaves)
        vals = dict(zip(_STATE_LEAVES, out[:len(_STATE_LEAVES)]))
        prims = []
        n0, npl = len(_STATE_LEAVES), len(_PRIM_LEAVES)
        for q, ps in enumerate(state.primitives):
            if q < n_prim:
                ps = ps._replace(**dict(zip(_PRIM_LEAVES, out[n0 + q * npl: n0 + (q + 1) * npl])))
            prims.append(ps)
        new_state = state._replace(primitives=prims, **vals)
        return new_state, new_state

    def sort_bins(self, x: torch.Tensor):
        """Exposes the per-frame binning: base [B,n,3], key [B,n], perm [B,n] (int32)."""
        x = _f32c(x)
        B, n = x.shape[0], x.shape[1]
        p = self.params(B=B, n=n)
        base = torch.empty((B, n, 3), dtype=torch.int32, device=x.device)
        key = torch.empty((B, n), dtype=torch.int32, device=x.device)
        perm = torch.empty((B, n), dtype=torch.int32, device=x.device)
        ws, nbytes = self._ws_fwd.get(self._L.ud_mpm_fwd_workspace_bytes(C.byref(p)))
        rc = self._L.ud_mpm_sort_bins(C.byref(p), _ptr(x), _ptr(base), _ptr(key), _ptr(perm), ws, nbytes,
                                      self._stream())
        _lib.check(rc, "ud_mpm_sort_bins")
        return base, key, perm
