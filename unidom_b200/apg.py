"""Analytic policy gradient (APG) around the B200 simulator step: the caller of `env.step_diff` and the ONLY
collective of the path.

Mirrors DaXBench/daxbench/algorithms/apg/apg.py:
  policy   make_direct_optimization_model (:353-358): MLP obs -> 512 -> 256 -> 2A, swish
  sampling brax NormalTanhDistribution: a = tanh(loc + (softplus(raw_scale) + 0.001) * eps); cloth envs apply
           sigmoid on top (:181-186)
  loss     -mean(rewards) over a lax.scan of ep_len env steps (:207-215)
  update   nan_to_num -> per-device global-norm clip -> pmean over devices -> adam (:233-240, 260-267)

Multi-GPU: one process per GPU; rank r simulates envs [r*B/N, (r+1)*B/N) (:83-85); the policy gradient lives in
ONE flat fp32 buffer that is scrubbed and clipped per rank and then averaged with a single all-reduce
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  The ORDER matters for parity: clip BEFORE the mean.
All tensor math here is device-agnostic torch (it is plumbing around the kernels, 0.9 M parameters).
"""
import math
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as Fnn


def shard_envs(num_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """apg.py:83-85: num_envs // devices per device; returns (first env, envs on this rank)."""
    if num_envs % world_size:
        raise ValueError(f"num_envs={num_envs} is not divisible by the number of devices ({world_size})")
    per = num_envs // world_size
    return rank * per, per


def init_policy(obs_size: int, action_size: int, seed: int = 0, device="cpu") -> List[torch.Tensor]:
    """[W1,b1,W2,b2,W3,b3] of the 512-256-2A MLP.  LeCun-uniform kernels / zero biases like flax's Dense default;
    the PRNG stream is torch's (brax's threefry init is not reproducible here), so weights are an INPUT of parity
    tests, never compared across frameworks."""
    g = torch.Generator().manual_seed(seed)
    sizes = [obs_size, 512, 256, 2 * action_size]
    params = []
    for fan_in, fan_out in zip(sizes[:-1], sizes[1:]):
        lim = math.sqrt(3.0 / fan_in)
        params.append(((torch.rand((fan_in, fan_out), generator=g) * 2 - 1) * lim).to(device))
        params.append(torch.zeros(fan_out, device=device))
    return params


def policy_apply(params: List[torch.Tensor], obs: torch.Tensor) -> torch.Tensor:
    h = obs
    for i in range(0, len(params), 2):
        h = h @ params[i] + params[i + 1]
        if i + 2 < len(params):
            h = Fnn.silu(h)           # linen.swish
    return h


def sample_actions(logits: torch.Tensor, eps: torch.Tensor, sigmoid: bool = True) -> torch.Tensor:
    """NormalTanhDistribution.sample (brax 0.0.13, min_std = 0.001) followed by apg.py:185-186."""
    loc, raw_scale = logits.chunk(2, dim=-1)
    a = torch.tanh(loc + (Fnn.softplus(raw_scale) + 0.001) * eps)
    return torch.sigmoid(a) if sigmoid else a


def apg_loss(env, params, state, eps: torch.Tensor, sigmoid: bool = True):
    """apg.py:177-215: roll the policy through `ep_len = eps.shape[0]` env steps; loss = -mean(reward)."""
    rewards = []
    for t in range(eps.shape[0]):
        obs = env.get_obs(state)
        actions = sample_actions(policy_apply(params, obs), eps[t], sigmoid)
        _, reward, _, info = env.step_diff(actions, state)
        state = info["state"]
        rewards.append(reward)
    rewards = torch.stack(rewards)
    return -rewards.mean(), rewards, state


def flatten(tensors) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for t in tensors])


def unflatten(flat: torch.Tensor, like) -> List[torch.Tensor]:
    out, o = [], 0
    for t in like:
        out.append(flat[o:o + t.numel()].view_as(t))
        o += t.numel()
    return out


def reduce_policy_gradient(flat_grad: torch.Tensor, max_grad_norm: float, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """apg.py:233-235 + :260-267 on one flat buffer, in the reference's order:
    nan_to_num (per rank) -> clip to max_grad_norm by the per-rank global norm -> mean over ranks.
    Returns (reduced gradient, this rank's raw gradient norm = the `grad_norm` metric of :242)."""
    if flat_grad.is_cuda:          # fused scrub + norm + clip (csrc/reward.cu); the collective below stays NCCL
        import ctypes as C

        from . import _lib
        g = flat_grad.detach().to(torch.float32).contiguous().clone()
        sumsq = torch.empty(1, dtype=torch.float32, device=g.device)
        st = C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
        _lib.check(_lib.lib().ud_apg_scrub_clip(C.c_void_p(g.data_ptr()), g.numel(), float(max_grad_norm),
                                                C.c_void_p(sumsq.data_ptr()), st), "ud_apg_scrub_clip")
        g_norm = torch.sqrt(sumsq[0])
    else:                          # host logic under the gloo CPU tests
        g = torch.nan_to_num(flat_grad)
        g_norm = torch.sqrt((g * g).sum())
        g = torch.where(g_norm < max_grad_norm, g, (g / g_norm) * max_grad_norm)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)       # one collective per iteration
        g = g / dist.get_world_size(group)
    return g, g_norm


class Adam:
    """optax.adam(lr) (b1 0.9, b2 0.999, eps 1e-8, eps_root 0) on the flat parameter buffer."""

    def __init__(self, n: int, lr: float, device):
        self.lr, self.b1, self.b2, self.eps = lr, 0.9, 0.999, 1e-8
        self.m = torch.zeros(n, device=device)
        self.v = torch.zeros(n, device=device)
        self.t = 0

    def step(self, flat_params: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
        self.t += 1
        if flat_params.is_cuda:    # one fused pass (csrc/reward.cu k_adam_step), same arithmetic order as below
            import ctypes as C

            from . import _lib
            p = flat_params.detach().to(torch.float32).contiguous().clone()
            g = g.detach().to(torch.float32).contiguous()
            ptr = lambda t: C.c_void_p(t.data_ptr())       # noqa: E731
            st = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
            _lib.check(_lib.lib().ud_adam_step(ptr(p), ptr(g), ptr(self.m), ptr(self.v), p.numel(), 1, self.lr, self.b1,
                                               self.b2, self.eps, self.t, st), "ud_adam_step")
            return p
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * g * g
        mhat = self.m / (1 - self.b1 ** self.t)
        vhat = self.v / (1 - self.b2 ** self.t)
        return flat_params - self.lr * mhat / (torch.sqrt(vhat) + self.eps)


def fused_update_available(device) -> bool:
    """The one-kernel update needs every rank's gradient buffer mapped into every peer (torch symmetric memory over
    NVLink) -- NCCL process group with more than one rank on one node."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return False
    if dist.get_backend() != "nccl" or not torch.device(device).type == "cuda":
        return False
    try:
        import importlib
        importlib.import_module("torch.distributed._symmetric_memory")
    except Exception:
        return False
    from . import _lib
    return hasattr(_lib.lib(), "ud_apg_fused_update")


class FusedUpdate:
    """scrub + per-rank clip -> mean over ranks -> Adam as ONE kernel per rank (csrc/apg_fused.cu): every rank stages its
    clipped gradient in a peer-mapped buffer, the ranks meet at a flag barrier in peer memory, and every rank then
    reads all N staged gradients over NVLink in rank order (so the replicas stay bit-identical), divides by N and
    applies Adam to its replica.  Same arithmetic, same order of operations as reduce_policy_gradient + Adam.step."""

    def __init__(self, n: int, lr: float, device, group=None):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self.n, self.lr, self.b1, self.b2, self.eps = n, lr, 0.9, 0.999, 1e-8
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.m = torch.zeros(n, device=device)
        self.v = torch.zeros(n, device=device)
        self.t = 0
        self._L = _lib.lib()
        # two staging slots (iteration parity): a rank may start staging iteration t+1 while a peer still reads t
        # + two slots of the reduced gradient (reduce-scatter + broadcast form)
        self.stage = symm_mem.empty(4 * n, dtype=torch.float32, device=self.device)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=self.device)
        self.flags.zero_()
        h_stage = symm_mem.rendezvous(self.stage, self.group.group_name)
        h_flags = symm_mem.rendezvous(self.flags, self.group.group_name)
        self._keep = (h_stage, h_flags)
        self.peer_stage = torch.tensor([int(p) for p in h_stage.buffer_ptrs], dtype=torch.int64, device=self.device)
        self.peer_flags = torch.tensor([int(p) for p in h_flags.buffer_ptrs], dtype=torch.int64, device=self.device)
        self.scratch = torch.zeros(8, dtype=torch.float32, device=self.device)
        dist.barrier(self.group)
        torch.cuda.synchronize(self.device)
        self._C = C

    def step(self, flat_params: torch.Tensor, flat_grad: torch.Tensor, max_grad_norm: float) -> torch.Tensor:
        C = self._C
        self.t += 1
        p = flat_params.detach().to(torch.float32).contiguous().clone()
        g = flat_grad.detach().to(torch.float32).contiguous()
        ptr = lambda t: C.c_void_p(t.data_ptr())       # noqa: E731
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        from . import _lib
        _lib.check(self._L.ud_apg_fused_update(ptr(p), ptr(g), ptr(self.m), ptr(self.v), self.n, float(max_grad_norm),
                                               self.lr, self.b1, self.b2, self.eps, self.t, self.rank, self.world,
                                               ptr(self.peer_stage), ptr(self.peer_flags), ptr(self.scratch), st),
                   "ud_apg_fused_update")
        return p


def train_iteration(env, params, opt: Adam, state, eps, max_grad_norm: float, sigmoid: bool = True, group=None):
    """One `minimize` call (apg.py:217-258) on this rank's env shard.  Returns (new params, metrics)."""
    req = [p.detach().clone().requires_grad_(True) for p in params]
    loss, rewards, _ = apg_loss(env, req, state, eps, sigmoid)
    grads = torch.autograd.grad(loss, req)
    g, g_norm = reduce_policy_gradient(flatten(grads), max_grad_norm, group)
    new_flat = opt.step(flatten([p.detach() for p in params]), g)
    new_params = unflatten(new_flat, params)
    return new_params, {"loss": float(loss.detach()), "reward": rewards.detach(), "grad_norm": float(g_norm)}


def para_stiffness(it: int, train_min: float, train_max: float) -> float:
    """apg_para.py:326-329: ONE cloth stiffness per training iteration, `np.random.seed(it);
    np.random.uniform(train_min_stiff, train_max_stiff)` -- NumPy's legacy global generator, reproduced bit for bit
    (every rank draws the same value, every env of the iteration shares it)."""
    import numpy as np
    return float(np.random.RandomState(it).uniform(train_min, train_max))


def main(argv=None):
    """`python -m unidom_b200.apg --env fold_cloth3 --ep_len 3 --num_envs 4 --lr 1e-4 --seed 0 --max_it 2`
    (the reference's README command, apg.py:384-443; one process per GPU under torchrun, envs sharded by rank)."""
    import argparse
    import os
    import time

    import numpy as np

    from . import confs, envs
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="fold_cloth3", choices=["fold_cloth1", "fold_cloth3", "fold_cloth1_para"])
    ap.add_argument("--ep_len", type=int, default=3)
    ap.add_argument("--num_envs", type=int, default=4)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--max_it", type=int, default=2)
    ap.add_argument("--max_grad_norm", type=float, default=0.3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--goal", default=None, help="goal point cloud (.npy, (Q,3)); the reference reads "
                    "core/envs/goals/<task>/goal.npy (cloth_env.py:45), which is a data file of the reference")
    # apg_para.py:540-563
    ap.add_argument("--train_min_stiff", type=int, default=200)
    ap.add_argument("--train_max_stiff", type=int, default=1800)
    ap.add_argument("--eval_min_stiff", type=int, default=100)
    ap.add_argument("--eval_max_stiff", type=int, default=2000)
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    first, per = shard_envs(args.num_envs, world, rank)
    para = args.env.endswith("_para")
    goal = np.load(args.goal).astype(np.float32) if args.goal else np.zeros((1, 3), np.float32)

    def make_env(it):
        if para:
            stiffness = para_stiffness(it, args.train_min_stiff, args.train_max_stiff)
            return envs.FoldCloth1ParaEnv(per, aux_reward=True, seed=args.seed, stiffness=stiffness, goal=goal, device=dev,
                                          eval_min_max_stiff=[args.eval_min_stiff, args.eval_max_stiff])
        conf = confs.ClothConf()
        return envs.ClothEnv(conf, per, 4, confs.fold_cloth_mask(conf), goal=goal, aux_reward=True, device=dev)

    env = make_env(0)
    params = init_policy(env.observation_size, env.action_size, seed=args.seed, device=dev)   # replicated
    opt = Adam(sum(p.numel() for p in params), args.lr, dev)
    g = torch.Generator().manual_seed(args.seed + 1)
    for it in range(args.max_it):
        if para and it:
            env = make_env(it)                            # apg_para.py:333-339: the env is rebuilt every iteration
        _, state = env.reset()
        eps = torch.randn((args.ep_len, args.num_envs, env.action_size), generator=g)[:, first:first + per].to(dev)
        t0 = time.perf_counter()
        params, m = train_iteration(env, params, opt, state, eps, args.max_grad_norm)
        torch.cuda.synchronize(dev)
        pn = float(torch.sqrt(sum((p * p).sum() for p in params)))
        extra = f" stiffness {float(state.stiffness[0]):.3f}" if para else ""
        print(f"[rank {rank}/{world}] it {it}: loss {m['loss']:.6f} grad_norm {m['grad_norm']:.4e} "
              f"params_norm {pn:.6f}{extra} ({time.perf_counter() - t0:.2f} s, envs {first}..{first + per - 1})", flush=True)
    if world > 1:
        # replicas must stay in sync: same reduced gradient + same Adam state on every rank
        flat = flatten(params)
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref), "policy replicas diverged"
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
