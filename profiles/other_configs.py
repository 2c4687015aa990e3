#!/usr/bin/env python
"""Throughput / memory of the BASELINE.json configs that are NOT the bench line (configs[2..4]), one GPU:
  pour_water   synthetic 99 998 liquid particles/env, 16 envs (= 128 envs over 8 GPUs), 2 bowl colliders, S=23
  whip_rope    long horizon: density 25 -> ~49 k particles, S=70 substeps/step, position control; peak HBM of an
               episode with step-granularity checkpoints vs the store-every-substep estimate
  fold_cloth1_para  128 envs/GPU (= 1024 over 8), float stiffness: cloth sub-action fwd+bwd
Prints one JSON line per config (run on the GPU box:  python profiles/other_configs.py > gpurun_out/other.jsonl)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unidom_b200 import _lib, confs  # noqa: E402
from unidom_b200.cloth_simulator import ClothSimulator  # noqa: E402
from unidom_b200.mpm_simulator import SimpleMPMSimulator, create_primitive  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def mpm_fwd_bwd(sim, state, action):
    leaves = {k: getattr(state, k).detach().requires_grad_(True) for k in ("x", "v", "C", "F")}
    a = action.detach().requires_grad_(True)
    out, _ = sim.step_jax(state._replace(**leaves), a)
    loss = (out.x * 1e-3).sum() + (out.v * 1e-4).sum()
    torch.autograd.grad(loss, list(leaves.values()) + [a])


def pour_water():
    B = 16
    conf = confs.pour_water_conf(res=(64, 48, 64))
    sim = SimpleMPMSimulator(conf, B, sdf_kind=_lib.UD_SDF_CONTAINER)
    st = sim.add_box(conf=conf, state=None, hardness=1, size=[0.3655] * 3, init_pos=[0.4, 0.3, 0.4], material=0, density=4)
    st.primitives.append(create_primitive(conf, 0.1, 666, [0.5] * 3, [0.35, 0.0, 0.02], [0.4, 0.3, 0.4]))
    st.primitives.append(create_primitive(conf, 0.1, 666, [0.5] * 3, [0.3, 0.0, 0.02], [0.4, 0.08, 0.2]))
    st = st._replace(primitives=[p._replace(action_scale=p.action_scale * 0.01) for p in st.primitives])
    state = sim.reset_jax(st)
    n = state.x.shape[1]
    action = torch.zeros((B, 12), device="cuda")
    action[:, 0], action[:, 5] = 0.3, 0.2
    with torch.no_grad():
        for _ in range(4):
            state, _ = sim.step_jax(state, action)
        fwd = timed(lambda: sim.step_jax(state, action))
    fb = timed(lambda: mpm_fwd_bwd(sim, state, action))
    units = B * n * conf.steps
    assert torch.isfinite(state.x).all()
    return {"config": "pour_water synthetic", "envs": B, "particles_per_env": n, "substeps": conf.steps, "res": list(conf.res),
            "fwd_ms": fwd, "fwdbwd_ms": fb, "fwd_Gpss": units / fwd / 1e6, "fwdbwd_Gpss": units / fb / 1e6,
            "particles_per_occupied_cell": float(n / len(torch.unique((state.x[0] * conf.inv_dx - 0.5).int(), dim=0)))}


def whip_rope():
    B, ep = 8, 5
    conf = confs.whip_rope_conf()
    sim = SimpleMPMSimulator(conf, B, use_position_control=True)
    st = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.15, 0.005, 0.005], init_pos=[0.25, 0.05, 0.25], material=1,
                     density=25)
    st.primitives.append(create_primitive(conf, 0.1, 666, [0.5] * 3, [0.02, 0.02, 0.02], [0.25, 0.05, 0.15]))
    st = st._replace(primitives=[p._replace(action_scale=p.action_scale * 0.02) for p in st.primitives])
    state = sim.reset_jax(st)
    n = state.x.shape[1]
    action = torch.zeros((B, 6), device="cuda")
    action[:, 1] = 0.5
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    t0 = time.perf_counter()
    x = state.x.detach().requires_grad_(True)
    a = action.detach().requires_grad_(True)
    s = state._replace(x=x)
    for _ in range(ep):
        s, _ = sim.step_jax(s, a)
    (gx, ga) = torch.autograd.grad((s.x * 1e-3).sum(), [x, a])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    peak = torch.cuda.max_memory_allocated() - base
    store_all = ep * conf.steps * B * n * 100 + ep * conf.steps * B * int(np.prod(conf.res)) * 32
    assert torch.isfinite(gx).all()
    return {"config": "whip_rope long horizon", "envs": B, "particles_per_env": n, "substeps_per_step": conf.steps,
            "episode_steps": ep, "episode_substeps": ep * conf.steps, "fwdbwd_s": dt,
            "fwdbwd_Gpss": B * n * conf.steps * ep / dt / 1e9, "peak_hbm_bytes": int(peak),
            "store_every_substep_bytes_estimate": int(store_all), "state_bytes_per_checkpoint": B * n * 100}


def cloth_para():
    B = 128
    conf = confs.ClothConf()
    sim = ClothSimulator(conf, B, None, confs.fold_cloth_mask(conf))
    st = sim.reset_jax()
    g = torch.Generator().manual_seed(0)
    st = st._replace(stiffness=(200 + 1600 * torch.rand(B, generator=g)).cuda(),
                     primitive0=torch.cat([st.x[:, 0], torch.full((B, 1), 0.01, device="cuda")], dim=1))
    act = torch.tensor([0.0, 0.3, 0.1, 0.0, 0, 0, 0, 1.0], device="cuda").repeat(B, 1)

    def fb():
        x = st.x.detach().requires_grad_(True)
        a = act.detach().requires_grad_(True)
        o, _ = sim.step_jax(st._replace(x=x), a)
        torch.autograd.grad((o.x * o.x).sum(), [x, a])
    with torch.no_grad():
        fwd = timed(lambda: sim.step_jax(st, act), n=20)
    t = timed(fb, n=20)
    # device time of the two kernels alone (CUDA events recorded by the library around its launches)
    import ctypes as C
    L = _lib.lib()
    L.ud_timing_enable(1)
    for _ in range(10):
        fb()
    torch.cuda.synchronize()
    ncls = L.ud_timing_num_classes()
    msb, cnt = (C.c_double * ncls)(), (C.c_int64 * ncls)()
    L.ud_timing_collect(msb, cnt, ncls)
    L.ud_timing_enable(0)
    L.ud_timing_class_name.restype = C.c_char_p
    kern = {L.ud_timing_class_name(i).decode(): msb[i] / cnt[i] for i in range(ncls) if cnt[i] > 0}
    units = B * sim.n_nodes * 50
    return {"config": "fold_cloth1_para sub-action", "kernel_ms": kern, "envs": B, "nodes": sim.n_nodes, "substeps": 50, "fwd_ms": fwd,
            "fwdbwd_ms": t, "fwd_Gnss": units / fwd / 1e6, "fwdbwd_Gnss": units / t / 1e6}


def cloth_env_step():
    """One ClothEnv.step_diff (40 sub-actions x 50 substeps, cloth_env.py:211) fwd+bwd: fused scan vs 40 calls."""
    from unidom_b200 import envs
    B = 128
    conf = confs.ClothConf()
    res = {"config": "fold_cloth3 env step (40 sub-actions x 50 substeps)", "envs": B}
    for fused in (True, False):
        env = envs.ClothEnv(conf, B, 4, confs.fold_cloth_mask(conf), goal=np.zeros((1, 3), np.float32), aux_reward=True,
                            fused=fused)
        _, state = env.reset()
        g = torch.Generator().manual_seed(0)
        act = torch.rand((B, env.action_size), generator=g).cuda()

        def fb():
            a = act.detach().requires_grad_(True)
            _, r, _, info = env.step_diff(a, state)
            torch.autograd.grad(r.sum(), [a])

        def fw():
            with torch.no_grad():
                env.step_diff(act, state)
        key = "fused" if fused else "stepwise"
        res[key + "_fwd_ms"] = timed(fw, n=5)
        res[key + "_fwdbwd_ms"] = timed(fb, n=5)
    units = B * 512 * 2000
    res["fused_fwdbwd_Gnss"] = units / res["fused_fwdbwd_ms"] / 1e6
    return res


def small_scene_graph():
    """The reference's own whip_rope size (67 particles, res 32^3, S = 70): launch-bound; eager vs CUDA-graph replay
    of a 5-step forward scan."""
    from unidom_b200.graphs import GraphedMPMScan
    B, T = 16, 5
    conf = confs.whip_rope_conf()
    sim = SimpleMPMSimulator(conf, B, use_position_control=True)
    st = sim.add_box(conf, None, size=[0.25, 0.01, 0.01], init_pos=[0.5, 0.02, 0.5], hardness=1.0, material=1,
                     density=2.75)
    prim = create_primitive(conf, friction=0.0, softness=666.0, color=[0.5] * 3, size=[0.01, 0.01, 0.01],
                            init_pos=[0.37, 0.02, 0.5])
    st = st._replace(primitives=[prim])
    state = sim.reset_jax(st)
    acts = torch.zeros((T, B, 6), device="cuda")
    acts[..., 1] = 0.5

    def eager():
        s = state
        with torch.no_grad():
            for t in range(T):
                s, _ = sim.step_jax(s, acts[t])
    graph = GraphedMPMScan(sim, state, acts)
    te = timed(eager, n=10)
    tg = timed(lambda: graph(state, acts), n=10)
    n = state.x.shape[1]
    return {"config": "whip_rope shipped size, forward scan", "envs": B, "particles_per_env": n, "substeps": conf.steps * T,
            "eager_ms": te, "graph_ms": tg, "eager_us_per_substep": 1e3 * te / (conf.steps * T),
            "graph_us_per_substep": 1e3 * tg / (conf.steps * T)}


def reward_kernels():
    """calc_chamfer fwd+bwd: fused kernels vs the reference's formulation (materialised (B,P,Q) distances) in torch."""
    from unidom_b200 import envs
    res = {"config": "calc_chamfer fwd+bwd"}
    for tag, B, P, torch_too in (("cloth_128x512x512", 128, 512, True), ("plasticine_32x50625x50625", 32, 50625, False)):
        g = torch.Generator().manual_seed(0)
        x = torch.rand((B, P, 3), generator=g).cuda()
        y = torch.rand((P, 3), generator=g).cuda()

        def fb(fn):
            xr = x.detach().requires_grad_(True)
            torch.autograd.grad(fn(xr, y).sum(), [xr])

        def ref(xr, yy):
            d = torch.sqrt(((xr[:, :, None, :] - yy[None, None, :, :]) ** 2).mean(-1))
            return d.amin(-1).mean(1) + d.amin(-2).mean(1)
        res[tag + "_ms"] = timed(lambda: fb(envs.calc_chamfer), n=5)
        res[tag + "_Gpairs_per_s"] = 3 * B * P * P / res[tag + "_ms"] / 1e6      # 2 scans + 1 adjoint pass
        if torch_too:
            res[tag + "_torch_ms"] = timed(lambda: fb(ref), n=5)
    return res


if __name__ == "__main__":
    only = sys.argv[1:] or ["pour_water", "whip_rope", "cloth_para", "cloth_env_step", "reward_kernels", "small_scene_graph"]
    for fn in [f for f in (pour_water, whip_rope, cloth_para, cloth_env_step, reward_kernels, small_scene_graph) if f.__name__ in only]:
        print(json.dumps(fn()), flush=True)
