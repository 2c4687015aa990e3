"""Host-side mirror of the reference's env wrappers around the simulator step (the callers of the hot path).

Mirrors DaXBench/daxbench/core/envs/basic/cloth_env.py (ClothEnv: get_obs :97-132, get_pnp_actions :136-173,
reset :178-188, step_diff :204-231) and core/utils/util.py (calc_chamfer :138-153, calc_l2 :156-159).
The rewards (calc_chamfer / calc_l2 and their adjoints) are kernels of libunidom_b200.so (csrc/reward.cu); the rest
is glue around `ClothSimulator.step_jax` / `SimpleMPMSimulator.step_jax`: torch ops on the simulator's device,
differentiable through torch autograd so that APG's rollout gradient reaches the policy.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .cloth_simulator import ClothSimulator, ClothState
from .mpm_simulator import _aligned, _ptr


def _dev_f32(t, like=None):
    t = torch.as_tensor(t)
    if like is not None and t.device != like.device:
        t = t.to(like.device)
    if not t.is_cuda:
        raise RuntimeError("unidom_b200 reward kernels need CUDA tensors (sm_100a); there is no CPU path")
    return t.detach().to(torch.float32).contiguous()


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _Chamfer(torch.autograd.Function):
    """ud_chamfer_fwd / ud_chamfer_bwd (csrc/reward.cu): no (B,P,Q,3) tensor, residuals = 8 B per point."""

    @staticmethod
    def forward(ctx, x, y):
        L = _lib.lib()
        xc, yc = _dev_f32(x), _dev_f32(y, x)
        B, P, Q = xc.shape[0], xc.shape[1], yc.shape[0]
        nbytes = L.ud_chamfer_residual_bytes(B, P, Q)
        res = torch.empty(nbytes + 256, dtype=torch.uint8, device=xc.device)
        out = torch.empty(B, dtype=torch.float32, device=xc.device)
        _lib.check(L.ud_chamfer_fwd(_ptr(xc), _ptr(yc), B, P, Q, _ptr(out), _aligned(res), nbytes, _stream(xc)),
                   "ud_chamfer_fwd")
        ctx.save_for_backward(xc, yc, res)
        return out

    @staticmethod
    def backward(ctx, g):
        xc, yc, res = ctx.saved_tensors
        L = _lib.lib()
        B, P, Q = xc.shape[0], xc.shape[1], yc.shape[0]
        gx = torch.empty_like(xc)
        _lib.check(L.ud_chamfer_bwd(_ptr(xc), _ptr(yc), B, P, Q, _ptr(_dev_f32(g)), _aligned(res), res.numel() - 256,
                                    _ptr(gx), _stream(xc)), "ud_chamfer_bwd")
        return gx, None


class _L2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        L = _lib.lib()
        xc = _dev_f32(x)
        yc = _dev_f32(y, x).expand(xc.shape[1], 3).contiguous()
        out = torch.empty(xc.shape[0], dtype=torch.float32, device=xc.device)
        _lib.check(L.ud_l2_fwd(_ptr(xc), _ptr(yc), xc.shape[0], xc.shape[1], _ptr(out), _stream(xc)), "ud_l2_fwd")
        ctx.save_for_backward(xc, yc)
        return out

    @staticmethod
    def backward(ctx, g):
        xc, yc = ctx.saved_tensors
        gx = torch.empty_like(xc)
        _lib.check(_lib.lib().ud_l2_bwd(_ptr(xc), _ptr(yc), xc.shape[0], xc.shape[1], _ptr(_dev_f32(g)), _ptr(gx),
                                        _stream(xc)), "ud_l2_bwd")
        return gx, None


def calc_chamfer(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """util.py:138-153.  x (B,P,3), y (Q,3) -> (B,).  The reference's point distance is sqrt(mean(d^2)) over the
    3 coordinates, not the Euclidean norm; ties of the minima share the cotangent like jnp.min's VJP.
    Differentiable in x only (the goal is a constant of the task)."""
    return _Chamfer.apply(x, y)


def calc_l2(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """util.py:156-159.  x (B,P,3), y (P,3) or (1,3) -> (B,)."""
    return _L2.apply(x, y)


def get_pnp_actions(actions: torch.Tensor, state: ClothState) -> torch.Tensor:
    """cloth_env.py:136-173, batched: pick-and-place (B,6) -> 40 sub-actions (40,B,8):
    3 approach (suction 1) + 10 lift + 20 move + 7 release; the second gripper's half is zero."""
    B, dev, dt = actions.shape[0], actions.device, actions.dtype
    zero = torch.zeros((B, 1), device=dev, dtype=dt)
    pick = torch.cat([actions[:, 0:1], zero, actions[:, 2:3]], dim=1)
    place = torch.cat([actions[:, 3:4], zero, actions[:, 5:6]], dim=1)
    act_down = torch.cat([(pick - state.primitive0[:, :3]) / 3, torch.ones((B, 1), device=dev, dtype=dt)], dim=1)
    act_down = act_down[None].expand(3, B, 4)
    up = torch.tensor([0.0, 0.06, 0.0], device=dev, dtype=dt) / 10
    act_up = torch.cat([up, torch.zeros(1, device=dev, dtype=dt)])[None, None].expand(10, B, 4)
    mv = place - pick
    mv = torch.cat([mv[:, 0:1], zero, mv[:, 2:3]], dim=1)
    act_move = torch.cat([mv / 20, zero], dim=1)[None].expand(20, B, 4)
    act_release = torch.tensor([0.0, 0.0, 0.0, 1.0], device=dev, dtype=dt)[None, None].expand(7, B, 4)
    sub = torch.cat([act_down, act_up, act_move, act_release], dim=0)
    return torch.cat([sub, torch.zeros_like(sub)], dim=2)


class ClothEnv:
    """B200 drop-in for the reference's ClothEnv (fold_cloth1/3, unfold_cloth1/3, fold_cloth1_para)."""

    def __init__(self, conf, batch_size, max_steps, cloth_mask, goal=None, aux_reward=False, device="cuda",
                 para=False, fused=True, obs_stride=1, eval_min_max_stiff=(100, 2000)):
        self.conf, self.batch_size, self.max_steps, self.aux_reward = conf, batch_size, max_steps, aux_reward
        self.fused = fused                                   # scan over the 40 sub-actions inside the library
        self.simulator = ClothSimulator(conf, batch_size, None, cloth_mask, device=device)
        self.device = self.simulator.device
        self.action_size = 6
        self.para = para                                     # cloth_env_para.py: obs carries the normalised stiffness
        self.eval_min_max_stiff = [float(eval_min_max_stiff[0]), float(eval_min_max_stiff[1])]   # cloth_env_para.py:41
        self.obs_stride = obs_stride                         # fold_cloth_tshirt_env.py:100: every 10th node
        n = self.simulator.n_nodes
        self.observation_size = -(-n // obs_stride) * 3 + 8 + (1 if para else 0)
        goal = np.zeros((1, 3), np.float32) if goal is None else np.asarray(goal, np.float32)
        self.goal = torch.from_numpy(goal).to(self.device)

    def get_obs(self, state: ClothState) -> torch.Tensor:
        """cloth_env.py:119-128 (PARTICLE): [x.flatten(), primitive0, primitive1]; the parameter-aware env appends
        (stiffness - eval_min) / (eval_max - eval_min) (cloth_env_para.py:124-131; the bounds are the env's
        `eval_min_max_stiff`, default [100, 2000], fold_cloth1_para_env.py:41)."""
        parts = [state.x[:, ::self.obs_stride].flatten(1), state.primitive0, state.primitive1]
        if self.para:
            lo, hi = self.eval_min_max_stiff
            parts.append(((state.stiffness.to(state.x.dtype) - lo) / (hi - lo))[:, None])
        return torch.cat(parts, dim=1)

    def reset(self, key=None, shift_xz=None):
        """cloth_env.py:178-188: `key, _ = split(key); x[..., [0, 2]] += normal(key, (2,)) * 0.05` -- the lattice plus ONE
        xz shift shared by all envs, drawn from the reference's own threefry stream (unidom_b200.jaxrng).  `key`: a
        PRNGKey (2 x uint32) or an int seed, default PRNGKey(0) as in apg.py; `shift_xz` (2,) overrides the draw."""
        st = self.simulator.reset_jax()
        if shift_xz is None:
            from . import jaxrng
            k = jaxrng.PRNGKey(0 if key is None else key) if (key is None or np.isscalar(key)) else np.asarray(key, np.uint32)
            k = jaxrng.split(k)[0]
            shift_xz = jaxrng.normal(k, (2,)) * np.float32(0.05)
        sh = torch.as_tensor(shift_xz, dtype=st.x.dtype, device=self.device)
        x = st.x.clone()
        x[..., 0] += sh[0]
        x[..., 2] += sh[1]
        st = st._replace(x=x)
        return self.get_obs(st), st

    def random_fold(self, state: ClothState, step=3, rng=None) -> ClothState:
        """unfold_cloth3_env.py:57-70: `step` pick-and-place actions between two random nodes of each env (the
        reference draws them from np.random), without gradients: the start state of the unfold tasks."""
        rng = rng or np.random
        B, P = state.x.shape[:2]
        bidx = torch.arange(B, device=self.device)
        for _ in range(step):
            st = torch.from_numpy(rng.randint(0, P, size=(B,))).to(self.device)
            ed = torch.from_numpy(rng.randint(0, P, size=(B,))).to(self.device)
            actions = torch.cat([state.x[bidx, st], state.x[bidx, ed]], dim=-1)
            with torch.no_grad():
                _, _, _, info = self.step_diff(actions, state)
            state = info["state"]
        return state

    def step_diff(self, actions: torch.Tensor, state: ClothState):
        """cloth_env.py:204-231: 40 sub-actions x 50 substeps through the simulator, reward
        e^(-10 chamfer) (+ e^(-contact)) * 0.99^cur_step."""
        old_chamfer = calc_chamfer(state.x, self.goal)
        contact = torch.sqrt(((actions[:, None, :3] - state.x) ** 2).sum(-1)).min(-1).values
        sub = get_pnp_actions(actions, state)
        if self.fused:                                  # lax.scan(self.simulator.step_jax, ...) (:211) as one call
            state = self.simulator.scan_step_jax(state, sub.contiguous())
        else:
            for a in sub:
                state, _ = self.simulator.step_jax(state, a)
        state = state._replace(cur_step=state.cur_step + 1)
        obs = self.get_obs(state)
        chamfer = calc_chamfer(state.x, self.goal)
        reward = math.e ** (-chamfer * 10)
        if self.aux_reward:
            reward = reward + math.e ** (-contact)
        reward = reward * 0.99 ** state.cur_step.to(reward.dtype)
        done = state.cur_step >= self.max_steps
        info = {"state": state, "real_reward": old_chamfer - chamfer + 0.1 * contact}
        return obs, reward, done, info


class FoldCloth1ParaEnv(ClothEnv):
    """core/envs/fold_cloth1_para_env.py:39-53 (GenDOM's parameter-aware fold task, BASELINE configs[3]): same
    constructor arguments; `stiffness` becomes `conf.stiffness` (a Python number: a float draw makes the state's
    stiffness leaf float32 and differentiable, the integer default 900 does not), the observation carries the
    stiffness normalised by `eval_min_max_stiff`; max_steps = 3, observation_size = 1545."""

    def __init__(self, batch_size, conf=None, aux_reward=False, seed=1, stiffness=900, eval_min_max_stiff=(100, 2000),
                 goal=None, device="cuda", fused=True):
        from . import confs
        conf = confs.FoldCloth1ParaConf() if conf is None else conf
        conf.stiffness = stiffness
        super().__init__(conf, batch_size, 3, confs.fold_cloth_mask(conf), goal=goal, aux_reward=aux_reward, device=device,
                         para=True, fused=fused, eval_min_max_stiff=eval_min_max_stiff)
        assert self.observation_size == 1545


class UnfoldClothEnv(ClothEnv):
    """core/envs/unfold_cloth1_env.py:40-82 / unfold_cloth3_env.py:40-83: the fold_cloth scene with friction mu = 3,
    max_steps = 15, observation_size 1544, whose reset is
        key, _ = split(key); x = lattice + normal(key, x.shape) * 1e-4; state = random_fold(state, step=n_folds)
    with n_folds = 1 (unfold_cloth1) or 3 (unfold_cloth3).  The noise comes from the reference's own threefry stream
    (unidom_b200.jaxrng), the fold end points from `rng` (the reference draws them from the global np.random state:
    pass np.random, or a RandomState seeded like it)."""

    def __init__(self, batch_size, conf=None, aux_reward=False, seed=1, n_folds=3, goal=None, device="cuda", fused=True):
        from . import confs
        conf = confs.UnfoldClothConf() if conf is None else conf
        super().__init__(conf, batch_size, 15, confs.fold_cloth_mask(conf), goal=goal, aux_reward=aux_reward,
                         device=device, fused=fused)
        self.n_folds = n_folds
        assert self.observation_size == 1544

    def reset(self, key=None, rng=None):
        from . import jaxrng
        st = self.simulator.reset_jax()
        k = jaxrng.PRNGKey(0 if key is None else key) if (key is None or np.isscalar(key)) else np.asarray(key, np.uint32)
        k = jaxrng.split(k)[0]
        noise = jaxrng.normal(k, tuple(st.x.shape)) * np.float32(0.0001)
        st = st._replace(x=st.x + torch.as_tensor(noise, dtype=st.x.dtype, device=self.device))
        self.reset_noisy_x = st.x.clone()                   # the state before the folds (tests compare it bit for bit)
        st = self.random_fold(st, step=self.n_folds, rng=rng)
        return self.get_obs(st), st


# -------------------------------------------------------------------------------------------------------------------
# MPM env (core/envs/basic/mpm_env.py) with the push task of core/envs/shape_elasto_plastic.py ("push_plasticine")
# -------------------------------------------------------------------------------------------------------------------
def _auto_reset_shifts(state, scale, device):
    """The per-env draw of the reference's vmapped auto_reset (whip_rope_env.py:96-99, pour_water_env.py:98-101):
    `key, _ = split(key); shift = normal(key, (2,)) * scale` on every env's own state.key.  Returns (shift [B,2] on
    `device`, the new keys [B,2] int32)."""
    from . import jaxrng
    keys = state.key.detach().cpu().numpy().view(np.uint32)
    new_keys = np.stack([jaxrng.split(k)[0] for k in keys])
    sh = np.stack([jaxrng.normal(k, (2,)) for k in new_keys]).astype(np.float32) * np.float32(scale)
    return torch.from_numpy(sh).to(device), torch.from_numpy(new_keys.view(np.int32).copy()).to(device)


class MPMEnv:
    """B200 drop-in for the reference's MPMEnv (mpm_env.py:18-167): focus shift (pre_step/post_step :99-125),
    task-specific `get_primitive_actions`, scan of `simulator.step_jax` over the sub-actions (:141), state
    nan_to_num (:150-154), reward e^(-10 l2) (+ e^(-contact)) (:91-94,156-158), auto-reset on done (:159-161)."""

    def __init__(self, conf, batch_size, max_steps, goal=None, aux_reward=False, focus_computation=True, device="cuda",
                 use_position_control=False, seed=0):
        from . import jaxrng
        from .mpm_simulator import SimpleMPMSimulator
        self.conf, self.batch_size, self.max_steps = conf, batch_size, max_steps
        self.aux_reward, self.focus_computation = aux_reward, focus_computation
        self.simulator = SimpleMPMSimulator(conf, batch_size, use_position_control, device=device)
        self.simulator.key_global = jaxrng.PRNGKey(seed)                             # mpm_env.py:54
        self.device = self.simulator.device
        self.action_size = 6
        goal = np.zeros((1, 3), np.float32) if goal is None else np.asarray(goal, np.float32)
        self.goal = torch.from_numpy(goal).to(self.device)
        self.init_state = None

    @staticmethod
    def get_obs(state):
        """mpm_env.py:60-76 (PARTICLE): [x.flatten(), v.flatten(), primitives[0].position.flatten()]."""
        return torch.cat([state.x.flatten(1), state.v.flatten(1), state.primitives[0].position.flatten(1)], dim=1)

    # task hooks (overridden per task)
    def get_primitive_actions(self, actions, state):
        raise NotImplementedError

    def process_pre_step_actions(self, actions, shift):
        raise NotImplementedError

    def auto_reset(self, init_state, state):
        return state

    def _shift(self, state, shift):
        prims = [p._replace(position=p.position + shift[:, None, :]) if q < self.conf.n_primitive else p
                 for q, p in enumerate(state.primitives)]
        return state._replace(x=state.x + shift[:, None, :], primitives=prims)

    def step_diff(self, actions, state):
        contact = torch.sqrt(((actions[:, None, :3] - state.x) ** 2).sum(-1)).min(-1).values
        shift = None
        if self.focus_computation:                                           # pre_step (:99-114)
            centre = state.x.mean(1)
            target = torch.tensor(self.conf.res, dtype=centre.dtype, device=centre.device) * 0.5 / self.conf.n_grid
            shift = target - centre
            shift = torch.cat([shift[:, 0:1], torch.zeros_like(shift[:, 1:2]), shift[:, 2:3]], dim=1)
            actions = self.process_pre_step_actions(actions, shift)
            state = self._shift(state, shift)
        sub, state = self.get_primitive_actions(actions, state)              # (B, T, 6 n_prim)
        for t in range(sub.shape[1]):                                        # lax.scan(self.simulator.step_jax, ...) (:141)
            state, _ = self.simulator.step_jax(state, sub[:, t])
        state = state._replace(cur_step=state.cur_step + 1)
        if self.focus_computation:                                           # post_step (:116-125)
            state = self._shift(state, -shift)
        done = state.cur_step >= self.max_steps
        state = state._replace(**{k: torch.nan_to_num(getattr(state, k)) for k in ("x", "v", "C", "F", "J")})
        reward = math.e ** (-calc_l2(state.x, self.goal) * 10)
        if self.aux_reward:
            reward = reward + math.e ** (-contact)
        if bool(done.any()) and self.init_state is not None:                 # auto_reset + where(done) (:159-161)
            new = self.auto_reset(self.init_state, state)
            state = _where_state(done, new, state)
        return self.get_obs(state), reward, done, {"state": state}


def _where_state(done, new, old):
    def pick(a, b):
        if not torch.is_tensor(a):
            return b
        d = done.reshape(done.shape + (1,) * (a.dim() - done.dim()))
        return torch.where(d, a.detach(), b)
    prims = [type(po)(*[pick(a, b) for a, b in zip(pn, po)]) for pn, po in zip(new.primitives, old.primitives)]
    vals = {k: pick(getattr(new, k), getattr(old, k)) for k in old._fields if k != "primitives"}
    return old._replace(primitives=prims, **vals)


class ShapeElastoPlasticEnv(MPMEnv):
    """core/envs/shape_elasto_plastic.py:56-157 (class ShapeRopeEnv there): a box pusher starts at `start`, moves
    towards `end` by at most 0.1 in 20 sub-actions (:95-123); BASELINE configs[1] "push_plasticine"."""

    MAX_PUSH, N_SUB = 0.1, 20            # :91, :99

    def __init__(self, conf, batch_size, max_steps=6, density=3.0, **kw):
        super().__init__(conf, batch_size, max_steps, focus_computation=True, **kw)
        self.state = self._build(density)
        self.init_state = self.state

    def _build(self, density):
        from . import confs
        return confs.build_shape_elasto_plastic(self.simulator, density=density)

    def process_pre_step_actions(self, actions, shift):
        return torch.cat([actions[:, 0:3] + shift, actions[:, 3:] + shift], dim=1)      # :88-92

    def get_primitive_actions(self, actions, state):
        start, end = actions[:, :3], actions[:, 3:]
        y = torch.full_like(start[:, 1:2], 0.01)
        start = torch.cat([start[:, 0:1], y, start[:, 2:3]], dim=1)
        end = torch.cat([end[:, 0:1], y, end[:, 2:3]], dim=1)
        norm = torch.sqrt(((end - start) ** 2).sum(-1, keepdim=True)) + 1e-8
        vec = (end - start) / norm
        end = start + vec * torch.minimum(torch.maximum(norm, torch.zeros_like(norm)), torch.full_like(norm, self.MAX_PUSH))
        p = state.primitives[0]
        position = torch.cat([start[:, None, :], p.position[:, 1:]], dim=1)
        prims = [p._replace(position=position)] + list(state.primitives[1:])
        push = (end - start)[:, None, :].expand(-1, self.N_SUB, -1) / self.N_SUB
        push = torch.cat([push[..., 0:1], torch.zeros_like(push[..., 1:2]), push[..., 2:3]], dim=-1)
        sub = torch.cat([push, torch.zeros_like(push)], dim=-1)
        return sub, state._replace(primitives=prims)


class ShapeRopeEnv(ShapeElastoPlasticEnv):
    """core/envs/shape_rope_env.py:70-174 (and shape_rope_hard_env.py, which only pushes the rope around 8 more times
    at reset): a thin plastic rope (582 particles) pushed by at most 0.3 in 30 sub-actions of conf.steps (133) substeps.
    The reference's reset ends with random pushes drawn from np.random (:124-131,173): call `random_push` for the same."""

    MAX_PUSH, N_SUB = 0.3, 30            # :103, :42,109

    def _build(self, density):
        from . import confs
        return confs.build_shape_rope(self.simulator, density=density)

    def auto_reset(self, init_state, state):
        return state                                                                 # :83-86 ("TODO" in the reference)

    def random_push(self, step=2, radius=0.05, rng=None):
        rng = rng or np.random
        for _ in range(step):
            pc = self.state.x[0].detach().cpu().numpy()
            ids = rng.randint(0, pc.shape[0], self.batch_size)
            ang = rng.random((self.batch_size,)) * np.pi * 2
            off = np.stack([np.cos(ang) * radius, np.zeros_like(ang), np.sin(ang) * radius], axis=1)
            acts = np.concatenate([pc[ids] - off, pc[ids] + off], axis=1).astype(np.float32)
            acts[:, 1] = 0
            with torch.no_grad():
                _, _, _, info = self.step_diff(torch.from_numpy(acts).to(self.device), self.state)
            self.state = info["state"]
        return self.state


class WhipRopeEnv(MPMEnv):
    """core/envs/whip_rope_env.py:78-137: an elastic rope whipped by a position-controlled box gripper; ONE sub-action
    of conf.steps (70) substeps per env step; BASELINE configs[4] "whip_rope long horizon" (max_steps 70)."""

    def __init__(self, conf, batch_size, max_steps=70, density=2.75, **kw):
        kw.setdefault("use_position_control", True)
        super().__init__(conf, batch_size, max_steps, focus_computation=True, **kw)
        from . import confs
        self.state = confs.build_whip_rope(self.simulator, density=density)
        self.init_state = self.state
        self._rng = np.random.RandomState(getattr(conf, "seed", 1))

    def process_pre_step_actions(self, actions, shift):
        return actions                                                               # :88-90

    def get_primitive_actions(self, actions, state):
        a = (actions + 1e-12) / 50.0                                                 # :107-113
        a = torch.cat([a[:, :3], torch.zeros_like(a[:, 3:])], dim=1)
        return a[:, None, :], state

    def auto_reset(self, init_state, state):
        """:92-104: the initial scene with a fresh N(0, 0.02^2) xz shift per env on rope and gripper, drawn from every
        env's own threefry key like the reference's vmapped auto_reset (state.key advances with it)."""
        sh, keys = _auto_reset_shifts(init_state, 0.02, self.device)
        shift = torch.stack([sh[:, 0], torch.zeros_like(sh[:, 0]), sh[:, 1]], dim=1)
        p = init_state.primitives[0]
        pos = p.position.clone()
        pos[:, 0] = pos[:, 0] + shift
        prims = [p._replace(position=pos)] + list(init_state.primitives[1:])
        return init_state._replace(x=init_state.x + shift[:, None, :], primitives=prims, key=keys)


class PourWaterEnv(MPMEnv):
    """core/envs/pour_water_env.py:66-137: 702 liquid particles in a bowl (cut hollow sphere, container SDF) that the
    6-vector action translates / tilts, a second bowl on the ground; one sub-action of conf.steps (23) substeps per env
    step; BASELINE configs[2] "pour_water".  `points`: the liquid's particle positions (the reference draws them
    with jax.random.uniform, mpm_simulator.py:89; default = a seeded NumPy draw of the same box)."""

    def __init__(self, conf, batch_size, max_steps=100, points=None, **kw):
        super().__init__(conf, batch_size, max_steps, focus_computation=True, **kw)
        from .mpm_simulator import create_primitive
        sim = self.simulator
        if points is None:
            state = sim.add_box(conf=conf, state=None, hardness=1, size=[0.07, 0.07, 0.07], init_pos=[0.5, 0.2, 0.5],
                                z_rotation_angle=0, material=0, density=4)
        else:
            state = sim.add_box_from_points(conf, None, torch.as_tensor(np.asarray(points), dtype=torch.float32),
                                            hardness=1, material=0)
        for size, pos in (([0.09, 0.0, 0.008], [0.5, 0.2, 0.5]), ([0.08, 0.0, 0.008], [0.5, 0.06, 0.3])):   # :123-129
            state.primitives.append(create_primitive(conf, friction=0.1, softness=666, color=[0.5, 0.5, 0.5], size=size,
                                                     init_pos=pos))
        self.state = sim.reset_jax(state)
        self.init_state = self.state
        self._rng = np.random.RandomState(getattr(conf, "seed", 1))

    def process_pre_step_actions(self, actions, shift):
        return actions                                                               # :92-94

    def get_primitive_actions(self, actions, state):
        a = torch.cat([actions[:, :3] / 500.0, actions[:, 3:6] / 500.0, torch.zeros_like(actions)], dim=1) + 1e-12   # :79-88
        a = torch.cat([a[:, 0:1], torch.zeros_like(a[:, 1:2]), a[:, 2:]], dim=1)
        return a[:, None, :], state

    def auto_reset(self, init_state, state):
        """:96-105: bowl 0 back to [0.5, 0.2, 0.5] plus an N(0, 0.02^2) xz shift per env from the env's own threefry key."""
        sh, keys = _auto_reset_shifts(init_state, 0.02, self.device)
        p = init_state.primitives[0]
        pos = p.position.clone()
        pos[:, 0] = torch.tensor([0.5, 0.2, 0.5], device=self.device) + torch.stack([sh[:, 0], torch.zeros_like(sh[:, 0]), sh[:, 1]], 1)
        return init_state._replace(primitives=[p._replace(position=pos)] + list(init_state.primitives[1:]), key=keys)
