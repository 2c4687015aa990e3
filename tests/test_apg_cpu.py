"""CPU (gloo, world_size 2): the N>1 host logic of the path -- env sharding and the one collective
(policy-gradient mean with the reference's ordering: nan_to_num -> per-rank clip -> mean, apg.py:233-235)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unidom_b200 import apg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_grad(rank, n):
    g = torch.Generator().manual_seed(100 + rank)
    v = torch.randn(n, generator=g)
    if rank == 0:
        v = v * 0.001            # norm below max_grad_norm: passes unclipped
        v[3] = float("nan")      # scrubbed to 0 before the norm
        v[5] = float("inf")      # scrubbed to FLT_MAX -> this rank ends up clipped after all
    else:
        v = v * 7.0              # clipped
    return v


def _worker(rank, world, port, n, max_norm, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, per = apg.shard_envs(8, world, rank)
    g, g_norm = apg.reduce_policy_gradient(_rank_grad(rank, n), max_norm)
    torch.save({"g": g, "norm": g_norm, "first": first, "per": per}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gradient_allreduce_order_and_env_sharding(tmp_path):
    n, max_norm, world = 1000, 0.3, 2
    mp.spawn(_worker, args=(world, _free_port(), n, max_norm, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    # expected: scrub and clip PER RANK, then mean
    exp = 0
    for r in range(world):
        g = torch.nan_to_num(_rank_grad(r, n))
        nrm = g.norm()
        exp = exp + (g if nrm < max_norm else g / nrm * max_norm)
        assert torch.allclose(res[r]["norm"], nrm)
    exp = exp / world
    for r in range(world):
        assert torch.allclose(res[r]["g"], exp, rtol=1e-6, atol=1e-9)      # identical on every rank
    # the other order (mean, then clip) gives a different vector: the test would catch a swapped implementation
    mean_first = sum(torch.nan_to_num(_rank_grad(r, n)) for r in range(world)) / world
    mean_first = mean_first / mean_first.norm() * max_norm
    assert (mean_first - exp).norm() > 0.05 * exp.norm()
    # env partition: contiguous, disjoint, complete
    assert [(res[r]["first"], res[r]["per"]) for r in range(world)] == [(0, 4), (4, 4)]
    with pytest.raises(ValueError):
        apg.shard_envs(7, 2, 0)


def test_adam_matches_torch_optim():
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(50, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt_ref = torch.optim.Adam([ref], lr=1e-3)
    opt = apg.Adam(50, 1e-3, "cpu")
    p = p0.clone()
    for _ in range(5):
        grad = torch.randn(50, generator=g)
        ref.grad = grad.clone()
        opt_ref.step()
        p = opt.step(p, grad)
    assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-7)


def test_policy_and_sampling_shapes():
    params = apg.init_policy(1544, 6, seed=0)
    assert sum(p.numel() for p in params) == 925452                    # the all-reduce message
    obs = torch.zeros((4, 1544))
    logits = apg.policy_apply(params, obs)
    a = apg.sample_actions(logits, torch.zeros((4, 6)))
    assert a.shape == (4, 6) and torch.allclose(a, torch.full_like(a, 0.5))   # sigmoid(tanh(0))
    flat = apg.flatten(params)
    back = apg.unflatten(flat, params)
    assert all(torch.equal(x, y) for x, y in zip(back, params))
