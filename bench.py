#!/usr/bin/env python
"""Benchmark of the DaXBench simulator step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config push_plasticine|pour_water|whip_rope|cloth_para] [--ckpt-window K]

The default (what the driver runs) is BASELINE configs[1]; --config selects the other BASELINE configs with the same
JSON line: pour_water (configs[2]: 99 998 liquid particles/env, two bowl colliders, 16 envs/GPU = 128 over 8),
whip_rope (configs[4]: 49 329 elastic particles/env, S = 70 substeps per step, adjoint with K-spaced substep
checkpoints, peak HBM reported), cloth_para (configs[3]: fold_cloth1_para, 128 envs/GPU = 1024 over 8, one env step
fwd+bwd + the policy-gradient all-reduce + Adam).

Workload (config.workload): BASELINE configs[1] "push_plasticine" = the reference's
envs/shape_elasto_plastic.py scene scaled to 50 625 particles/env (add_box density 3.9),
num_envs = 32 per GPU, S = 16 substeps per step, forward + backward (adjoint with
recompute) of one `SimpleMPMSimulator.step_jax` call = one "step".

metric  = particle-substeps/s, forward+backward:  envs * particles * substeps / time.
value   = inputs resident in HBM, CUDA events around exactly K steps, max over ranks.
e2e     = same metric through the public API with HOST (pinned) buffers: H2D of the step inputs,
          fwd+bwd, D2H of the new state and all gradients inside the timed region.
roofline= dominant kernel class: algorithmic bytes per launch / mean launch time (CUDA events on
          the launching stream, measured live) against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline = the CPU oracle (torch restatement of the reference; JAX is not installable here)
          timed on the host cores on a bounded sample of the same scene.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# algorithmic bytes per particle per launch of each kernel class:
ALG_BYTES = {"p2g": 132, "g2p": 72, "g2p_bwd": 84, "p2g_bwd": 240}
ALG_BYTES_FWD, ALG_BYTES_FWDBWD = 200, 700          # SURVEY.md section 8(d) contract figures
ALG_BYTES_TAPED = 500                               # section 8(d) "store-all variant": no recompute pass

DENSITY, ENVS_PER_GPU = 3.9, 32


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_scene(sim, density):
    """shape_elasto_plastic reset (envs/shape_elasto_plastic.py:139-157) already shifted to the
    focus point the env's pre_step uses (mpm_env.py:99-114): centroid at res/2 cells."""
    from unidom_b200.mpm_simulator import create_primitive
    conf = sim.conf
    state = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.2, 0.06, 0.12], init_pos=[0.25, 0.07, 0.25],
                        z_rotation_angle=0, material=2, density=density)
    state.primitives.append(create_primitive(conf, friction=0.1, softness=666, color=[0.5, 0.5, 0.5],
                                             size=[0.015, 0.06, 0.015], init_pos=[0.25, 0.01, 0.20]))
    return sim.reset_jax(state)


def build_pour_water(sim):
    """BASELINE configs[2]: pour_water scaled to ~100 k liquid particles per env (envs/pour_water_env.py:92-123 scene:
    a liquid box and two bowls with the container SDF)."""
    from unidom_b200.mpm_simulator import create_primitive
    conf = sim.conf
    st = sim.add_box(conf=conf, state=None, hardness=1, size=[0.3655] * 3, init_pos=[0.4, 0.3, 0.4], material=0, density=4)
    st.primitives.append(create_primitive(conf, 0.1, 666, [0.5] * 3, [0.35, 0.0, 0.02], [0.4, 0.3, 0.4]))
    st.primitives.append(create_primitive(conf, 0.1, 666, [0.5] * 3, [0.3, 0.0, 0.02], [0.4, 0.08, 0.2]))
    st = st._replace(primitives=[p._replace(action_scale=p.action_scale * 0.01) for p in st.primitives])
    return sim.reset_jax(st)


def build_whip_rope(sim):
    """BASELINE configs[4]: the whip_rope scene (envs/whip_rope_env.py:27-73, confs.build_whip_rope) at add_box density 25
    = 49 329 elastic particles, position-controlled gripper, S = 70 substeps per step."""
    from unidom_b200 import confs
    st = confs.build_whip_rope(sim, density=25)
    return st._replace(primitives=[p._replace(action_scale=p.action_scale * 0.02) for p in st.primitives])


def make_workload(args, dev, rank):
    """-> (sim, state, action, description) of the selected BASELINE config."""
    from unidom_b200 import _lib, confs
    from unidom_b200.mpm_simulator import SimpleMPMSimulator
    g = torch.Generator().manual_seed(1234 + rank)
    if args.config == "push_plasticine":
        conf = confs.shape_elasto_plastic_conf()
        sim = SimpleMPMSimulator(conf, args.envs, device=dev, p2g_mode=args.p2g_mode, adjoint=args.adjoint,
                                 ckpt_window=args.ckpt_window, env_groups=args.env_groups)
        state = build_scene(sim, args.density)
        B = state.x.shape[0]
        action = (torch.tensor([0.003, 0.0, 0.004, 0.0, 0.0, 0.0]) + 5e-4 * torch.randn((B, 6), generator=g)).to(dev)
        action[:, 3:] = 0
        name = workload_name(state.x.shape[1], B)
    elif args.config == "pour_water":
        conf = confs.pour_water_conf(res=(64, 48, 64))
        B = args.envs if args.envs != ENVS_PER_GPU else 16
        sim = SimpleMPMSimulator(conf, B, device=dev, sdf_kind=_lib.UD_SDF_CONTAINER, p2g_mode=args.p2g_mode,
                                 adjoint=args.adjoint, ckpt_window=args.ckpt_window, env_groups=args.env_groups)
        state = build_pour_water(sim)
        action = torch.zeros((B, 12), device=dev)
        action[:, 0], action[:, 5] = 0.3, 0.2
        action += (2e-2 * torch.randn((B, 12), generator=g)).to(dev)
        name = (f"pour_water MLS-MPM liquid fwd+bwd, {state.x.shape[1]} particles/env, num_envs={B}/GPU, "
                f"res={tuple(conf.res)}, {conf.steps} substeps/step, two container colliders")
    elif args.config == "whip_rope":
        conf = confs.whip_rope_conf()
        B = args.envs if args.envs != ENVS_PER_GPU else 8
        sim = SimpleMPMSimulator(conf, B, use_position_control=True, device=dev, p2g_mode=args.p2g_mode,
                                 adjoint=args.adjoint, ckpt_window=args.ckpt_window if args.ckpt_window else 10,
                                 env_groups=args.env_groups)
        state = build_whip_rope(sim)
        action = torch.tensor([[0.2, 0.5, 0.1, 0.0, 0.0, 0.0]], device=dev).repeat(B, 1)
        action[:, :3] += (2e-2 * torch.randn((B, 3), generator=g)).to(dev)
        name = (f"whip_rope MLS-MPM elastic fwd+bwd, {state.x.shape[1]} particles/env, num_envs={B}/GPU, "
                f"res={tuple(conf.res)}, {conf.steps} substeps/step, position control, "
                f"substep checkpoints every {sim.ckpt_window}")
    else:
        raise ValueError(args.config)
    return sim, state, action, name


def bind_to_local_numa(local_rank):
    """Pin this rank's host threads (and therefore the first-touch placement of its pinned buffers) to the CPUs NVML
    reports as local to its GPU.  Returns a description for the bench line."""
    info = {"cpus": None, "numa_nodes": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = f"{cpus[0]}-{cpus[-1]} ({len(cpus)})"
    except Exception as e:      # noqa: BLE001 -- affinity is an optimisation, never a requirement
        info["error"] = repr(e)[:80]
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        info["numa_nodes"] = len(nodes)
    except Exception:
        pass
    return info


def apg_update_leg(dev, dist, world, iters=20):
    """The path's one collective (algorithms/apg/apg.py:233-240, 260-267) on the policy gradient of the 512-256 MLP
    (925 964 fp32): ud_apg_scrub_clip -> all-reduce -> ud_adam_step, timed on the device inside the region."""
    from unidom_b200 import apg
    n = 925964
    g = torch.Generator().manual_seed(7)
    grad = (torch.randn(n, generator=g) * 1e-3).to(dev)
    params = torch.randn(n, generator=g).to(dev)
    opt = apg.Adam(n, 1e-4, dev)
    out = {}
    from unidom_b200 import _lib
    # fused: all-read form (every rank reads all N staged gradients) and reduce-scatter + broadcast form, both timed;
    # `fused_us` = the form the library picks for this world size (reduce-scatter from 4 ranks up)
    for label, fused, rs in (("nccl", False, -1), ("fused_allread", True, 0), ("fused_rs", True, 1)):
        if fused and not apg.fused_update_available(dev):
            continue
        _lib.lib().ud_tuning_set(b"apg_rs", rs)
        try:
            upd = apg.FusedUpdate(n, 1e-4, dev) if fused else None
        except Exception as e:      # noqa: BLE001 -- e.g. no peer mapping between some pair of GPUs: keep the NCCL number
            out["fused_error"] = repr(e)[:160]
            continue
        p = params.clone()
        for _ in range(3):
            p = upd.step(p, grad, 0.3) if fused else opt.step(p, apg.reduce_policy_gradient(grad, 0.3)[0])
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            p = upd.step(p, grad, 0.3) if fused else opt.step(p, apg.reduce_policy_gradient(grad, 0.3)[0])
        e1.record()
        dist.barrier()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[label + "_us"] = float(t.item())
    _lib.lib().ud_tuning_set(b"apg_rs", -1)
    pick = "fused_rs_us" if world >= 4 else "fused_allread_us"
    if pick in out:
        out["fused_us"] = out[pick]
    out.update({"elements": n, "bytes_reduced_per_rank": 4 * n, "ranks": world,
                "what": "scrub + per-rank global-norm clip -> mean over ranks -> Adam, per update, max over ranks"})
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU through NVML every 10 ms DURING the timed region
    (nvidia-smi takes longer per query than the whole region; it is the fallback when pynvml is missing)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        try:
            pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        except Exception:
            pw = 0.0
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        return [str(sm), str(mx)] + [("Active" if r & bits[k] else "Not Active") for k in
                                     ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")] + [pw]

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([c.strip() for c in out.strip().split(",")] + [0.0])
            except Exception:
                pass
            self._stop_evt.wait(0.01 if self.nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = sorted(int(r[0]) for r in self.rows if r and str(r[0]).isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and str(r[1]).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if str(v).lower().startswith("active")})
        pw = [float(r[6]) for r in self.rows if len(r) > 6]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": round(max(pw), 1) if pw else None,
                "reasons": reasons, "samples": len(sm), "via": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_cotangents(state, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    B, n = state.x.shape[:2]
    dev = state.x.device
    return {"x": (torch.randn((B, n, 3), generator=g) * 1e-3).to(dev),
            "v": (torch.randn((B, n, 3), generator=g) * 1e-4).to(dev)}


def fwd_bwd(sim, state, action, cot):
    """One step through the public API: forward (S substeps) + adjoint w.r.t. state and action."""
    leaves = {k: getattr(state, k).detach().requires_grad_(True) for k in ("x", "v", "C", "F")}
    a = action.detach().requires_grad_(True)
    out, _ = sim.step_jax(state._replace(**leaves), a)
    loss = (out.x * cot["x"]).sum() + (out.v * cot["v"]).sum()
    grads = torch.autograd.grad(loss, list(leaves.values()) + [a])
    return out, grads, loss


def detach_state(s):
    from unidom_b200.mpm_simulator import PrimitiveState
    prims = [PrimitiveState(*[t.detach() for t in p]) for p in s.primitives]
    vals = {k: getattr(s, k).detach() for k in s._fields if k != "primitives"}
    return s._replace(primitives=prims, **vals)


def cpu_oracle_rate(conf, n_envs, density, substeps, repeat, threads):
    """particle-substeps/s (fwd+bwd) of the CPU oracle on a bounded sample of the same scene."""
    import numpy as np
    from oracle import mpm as omp, primitives as oP
    torch.set_num_threads(threads)
    oconf = omp.MPMConf(n_grid=conf.n_grid, res=tuple(conf.res), dt=conf.dt, steps=substeps, E=conf.E, nu=conf.nu,
                        ground_friction=conf.ground_friction, gravity=tuple(conf.gravity),
                        n_primitive=conf.n_primitive, sdf_kind=conf.sdf_kind)
    x = omp.add_box(oconf, [0.2, 0.06, 0.12], [0.25, 0.07, 0.25], density=density)
    n = x.shape[0]
    prim = oP.create_primitive(substeps, 0.1, 666.0, [0.5] * 3, [0.015, 0.06, 0.015], [0.25, 0.01, 0.20])
    sim = omp.Simulator(oconf, torch.full((n,), 2, dtype=torch.int32), torch.ones(n))
    st = omp.reset_state(oconf, x, [prim], n_envs)
    act = torch.tensor([[0.003, 0.0, 0.004, 0.0, 0.0, 0.0]]).repeat(n_envs, 1)
    g = torch.Generator().manual_seed(0)
    cx = torch.randn((n_envs, n, 3), generator=g) * 1e-3
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        xs = st.x.clone().requires_grad_(True)
        a = act.clone().requires_grad_(True)
        out = omp.step_batch(sim, st._replace(x=xs), a)
        ((out.x * cx).sum()).backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_envs * n * substeps / best, n, best


def shim_reference_rate(density, substeps, repeat, threads):
    """particle-substeps/s (fwd+bwd) of the UNMODIFIED reference sources -- daxbench.core.engine.mpm_simulator.
    SimpleMPMSimulator.step_jax and jax.grad through its own custom_vjp rules -- executed under oracle/jaxshim (a
    torch-CPU stand-in for the jax API the path uses; JAX itself is not installable here), on the same bounded sample as
    cpu_oracle_rate.  Only where the reference tree exists (this container: /root/reference does not travel to the
    GPU box).  Returns (rate, n, seconds) or None."""
    ref = os.environ.get("UNIDOM_REFERENCE", "/root/reference/DaXBench")
    if not os.path.isdir(ref):
        return None
    import numpy as np
    torch.set_num_threads(threads)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "oracle"))
    import gen_golden as gg
    mods = gg.load_reference(False)
    import jax
    import jax.numpy as jnp
    mpm, _, prim, box, _ = mods
    conf = gg.MPMConf(jnp, jax.random.PRNGKey(0), n_grid=96, res=(48, 32, 48), dt=2e-4, steps=substeps, E=2, nu=0.2,
                      ground_friction=2, n_primitive=1)
    prim.set_sdf(box._sdf_batch)
    sim = mpm.SimpleMPMSimulator(conf, 1, use_position_control=False)
    sim.key_global = jax.random.PRNGKey(1)
    state = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.2, 0.06, 0.12], init_pos=[0.25, 0.07, 0.25],
                        z_rotation_angle=0, material=2, density=density)
    p = prim.create_primitive(conf, friction=0.1, softness=666, color=[0.5] * 3, size=[0.015, 0.06, 0.015],
                              init_pos=[0.25, 0.01, 0.20])
    state = sim.reset_jax(state._replace(primitives=[p]))
    n = state.x.shape[1]
    action = jnp.array(np.array([[0.003, 0.0, 0.004, 0.0, 0.0, 0.0]], np.float32))
    cx = jnp.array((np.random.RandomState(0).randn(1, n, 3) * 1e-3).astype(np.float32))

    def loss(inp):
        st, act = inp
        ns, _ = sim.step_jax(st, act)
        return (ns.x * cx).sum()
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        jax.grad(loss, allow_int=True)((state, action))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n * substeps / best, n, best


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  The reference itself (JAX) cannot
    be installed in this image (no jax/jaxlib wheel, no network), so this times the oracle PORT
    (oracle/, a torch-CPU restatement of the same arithmetic) with all host threads.  `--ref-kind shim` (only where
    /root/reference exists, i.e. in the build container, not on the GPU box) times the unmodified reference sources
    under oracle/jaxshim instead and reports the port beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_kind == "shim":
        threads = os.cpu_count() or 1
        sub = 2
        got = shim_reference_rate(DENSITY, sub, max(args.steps, 1), threads)
        if got is not None:
            rate, n, secs = got
            from unidom_b200 import confs
            port, _, _ = cpu_oracle_rate(confs.shape_elasto_plastic_conf(), 1, DENSITY, sub, 2, threads)
            line = {
                "impl": "reference", "metric": "particle-substeps/s fwd+bwd", "value": rate, "unit": "particle-substeps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(n), "sample": f"1 env x {n} particles x {sub} substeps fwd+bwd per step"},
                "cpu_baseline": {"value": rate, "unit": "particle-substeps/s", "cores": threads, "kind": "shim",
                                 "sample": f"1 env x {n} particles x {sub} substeps fwd+bwd (best of {max(args.steps, 1)}), the "
                                           "unmodified reference SimpleMPMSimulator.step_jax + jax.grad under oracle/jaxshim "
                                           "(torch-CPU)", "port_value": port},
                "e2e": {"value": rate, "unit": "particle-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            }
            print(json.dumps(line))
            return
    from unidom_b200 import confs
    conf = confs.shape_elasto_plastic_conf()
    threads = os.cpu_count() or 1
    sub = 2
    for _ in range(max(args.warmup, 0)):
        cpu_oracle_rate(conf, 1, 1.5, sub, 1, threads)     # short warm-up on a lighter sample
    rates, secs = [], []
    for _ in range(args.steps):
        r, n, dt = cpu_oracle_rate(conf, 1, DENSITY, sub, 1, threads)
        rates.append(r)
        secs.append(dt)
    total = sum(secs)
    value = len(secs) * n * sub / total
    line = {
        "impl": "reference", "metric": "particle-substeps/s fwd+bwd", "value": value, "unit": "particle-substeps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(n), "sample": f"1 env x {n} particles x {sub} substeps fwd+bwd per step"},
        "cpu_baseline": {"value": value, "unit": "particle-substeps/s", "cores": threads, "kind": "port",
                         "sample": f"1 env x {n} particles x {sub} substeps fwd+bwd, torch-CPU oracle"},
        "e2e": {"value": value, "unit": "particle-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_cloth_para(args, dev, dist, rank, world):
    """BASELINE configs[3]: fold_cloth1_para parameter-aware APG (GenDOM), 128 envs per GPU (= 1024 over 8).  A "step"
    is one APG training iteration with ep_len 1 on this rank's env shard: policy MLP, FoldCloth1ParaEnv.step_diff
    (40 pick-and-place sub-actions x 50 substeps, ONE fused forward launch + adjoint), policy gradient, per-rank
    scrub/clip, mean over ranks (the path's one collective), Adam.  Stiffness: one draw per iteration
    (apg_para.py:326-329).  value = cloth node-substeps/s over all ranks."""
    import numpy as np
    from unidom_b200 import apg, envs
    per = args.envs if args.envs != ENVS_PER_GPU else 128
    goal = np.zeros((1, 3), np.float32)
    env = envs.FoldCloth1ParaEnv(per, aux_reward=True, seed=0, stiffness=apg.para_stiffness(0, 200, 1800), goal=goal,
                                 device=dev, eval_min_max_stiff=[100, 2000])
    params = apg.init_policy(env.observation_size, env.action_size, seed=0, device=dev)
    n_par = sum(p.numel() for p in params)
    opt = apg.Adam(n_par, 1e-4, dev)
    g = torch.Generator().manual_seed(1 + rank)
    _, state = env.reset()
    eps = torch.randn((1, per, env.action_size), generator=g).to(dev)
    nodes = state.x.shape[1]
    units = per * nodes * 40 * 50

    def it():
        nonlocal params
        params, m = apg.train_iteration(env, params, opt, state, eps, 0.3)
        return m
    for _ in range(max(args.warmup, 3)):
        it()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m = it()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    apg_leg = apg_update_leg(dev, dist, world) if dist is not None else None
    if rank == 0:
        print(json.dumps({
            "metric": "node-substeps/s fwd+bwd", "value": world * units * args.steps / (ms * 1e-3), "unit": "node-substeps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"fold_cloth1_para parameter-aware APG iteration (ep_len 1), {per} envs/GPU, {nodes} nodes, "
                                   "40 sub-actions x 50 substeps per env step, policy 1545-512-256-16",
                       "envs_per_gpu": per, "policy_parameters": n_par,
                       "collective": "policy-gradient mean over ranks (one all-reduce per iteration)" if world > 1 else "none (1 GPU)"},
            "loss": m["loss"], "grad_norm": m["grad_norm"], "apg_update": apg_leg, "clocks": clocks}))
    if dist is not None:
        dist.destroy_process_group()


def workload_name(n, envs=ENVS_PER_GPU):
    return (f"push_plasticine = shape_elasto_plastic MLS-MPM fwd+bwd, {n} particles/env, "
            f"num_envs={envs}/GPU, res=(48,32,48), 16 substeps/step")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-kind", default="port", choices=["port", "shim"],
                    help="--impl reference: the oracle port (default; the only one available on the GPU box) or the "
                         "unmodified reference sources under oracle/jaxshim (where /root/reference exists)")
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--density", type=float, default=DENSITY)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--p2g-mode", type=int, default=0)
    ap.add_argument("--tune", default="", help="development A/B switches, e.g. svd_warm=0")
    ap.add_argument("--adjoint", default="recompute", choices=["auto", "tape", "recompute"],
                    help="recompute (headline, what north_star asks for): keep the step input only and re-run the "
                         "substeps in the backward; tape: the forward keeps its substep residuals in HBM (reported "
                         "next to the headline as `taped`)")
    ap.add_argument("--settle", type=int, default=8, help="env steps run before timing to reach a mid-push state")
    ap.add_argument("--no-e2e", action="store_true", help="development: skip the host-buffer leg")
    ap.add_argument("--config", default="push_plasticine", choices=["push_plasticine", "pour_water", "whip_rope", "cloth_para"],
                    help="BASELINE.json config to time (the driver's line is the default)")
    ap.add_argument("--env-groups", type=int, default=2,
                    help="run the envs of a call as this many sub-batches on separate CUDA streams (SimpleMPMSimulator)")
    ap.add_argument("--ckpt-window", type=int, default=None,
                    help="K-spaced substep checkpoints inside the recompute adjoint (default: none; whip_rope: 10)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_local_numa(local_rank)          # before any pinned allocation: first touch places the pages
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from unidom_b200 import _lib
    L = _lib.lib()
    for kv in filter(None, args.tune.split(",")):
        name, val = kv.split("=")
        assert L.ud_tuning_set(name.encode(), int(val)) >= 0, kv
    if args.config == "cloth_para":
        return run_cloth_para(args, dev, dist, rank, world)
    sim, state, action, wl_name = make_workload(args, dev, rank)
    conf = sim.conf
    B, n = state.x.shape[:2]
    S = conf.steps
    cot = make_cotangents(state, 99 + rank)
    units_per_step = B * n * S

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- settle: advance the scene so that every timed step starts from the SAME representative
    # mid-push state (deformed plasticine, F != I); otherwise ms/step would depend on how many steps ran before
    with torch.no_grad():
        for _ in range(args.settle):
            state, _ = sim.step_jax(state, action)
    state = detach_state(state)

    # ---------------- device-resident timing (value)
    for _ in range(max(args.warmup, 3)):
        out, grads, _ = fwd_bwd(sim, state, action, cot)
    barrier()
    torch.cuda.reset_peak_memory_stats(dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    L.ud_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out, grads, _ = fwd_bwd(sim, state, action, cot)
    e1.record()
    barrier()
    launches = int(L.ud_launch_count(0))
    clocks = sampler.stop()
    peak_hbm = int(torch.cuda.max_memory_allocated(dev))
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * units_per_step * args.steps / (ms * 1e-3)

    # ---------------- forward-only rate (reported next to the headline)
    with torch.no_grad():
        for _ in range(2):
            sim.step_jax(state, action)
        torch.cuda.synchronize(dev)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            sim.step_jax(state, action)
        f1.record()
        torch.cuda.synchronize(dev)
    fwd_value = world * units_per_step * args.steps / (f0.elapsed_time(f1) * 1e-3)

    # ---------------- the same step with the taped adjoint (forward keeps the substep residuals; no recompute pass)
    taped = None
    if args.adjoint == "recompute":
        sim.adjoint = "tape"
        for _ in range(max(args.warmup, 3)):        # same reference-holding pattern as the timed loop: the allocator
            out_t, grads_t, _ = fwd_bwd(sim, state, action, cot)   # must not meet new peak sizes inside it
        torch.cuda.synchronize(dev)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            out_t, grads_t, _ = fwd_bwd(sim, state, action, cot)
        t1.record()
        barrier()
        tms = t0.elapsed_time(t1)
        if dist is not None:
            t = torch.tensor([tms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tms = float(t.item())
        tv = world * units_per_step * args.steps / (tms * 1e-3)
        taped = {"value": tv, "unit": "particle-substeps/s", "ms_per_step": tms / args.steps,
                 "tape_bytes_per_step": int(L.ud_mpm_tape_bytes(C.byref(sim.params()))),
                 "algorithmic_bytes_per_particle_substep": ALG_BYTES_TAPED,
                 "step_frac": (tv / world) * ALG_BYTES_TAPED / 1e9 / peaks()[0]}
        del out_t, grads_t
        sim.adjoint = "recompute"

    # ---------------- per-kernel-class timing (roofline), live CUDA events on the launch stream
    groups_used = sim.env_groups
    sim.env_groups = 1             # per-kernel durations are taken with the envs as ONE batch on one stream
    for _ in range(2):
        fwd_bwd(sim, state, action, cot)
    L.ud_timing_enable(1)
    nprof = min(args.steps, 5)
    for _ in range(nprof):
        out, grads, _ = fwd_bwd(sim, state, action, cot)
    torch.cuda.synchronize(dev)
    ncls = L.ud_timing_num_classes()
    msb = (C.c_double * ncls)()
    cnt = (C.c_int64 * ncls)()
    L.ud_timing_collect(msb, cnt, ncls)
    L.ud_timing_enable(0)
    sim.env_groups = groups_used
    L.ud_timing_class_name.restype = C.c_char_p
    classes = {L.ud_timing_class_name(i).decode(): (msb[i], cnt[i]) for i in range(ncls) if cnt[i] > 0}
    tot_ms = sum(v[0] for v in classes.values())
    peak, peak_src = peaks()
    kern = {}
    for name, (m, c) in classes.items():
        avg_ms = m / c
        ent = {"share": m / tot_ms, "avg_ms": avg_ms, "launches_per_step": c / nprof}
        if name in ALG_BYTES:
            ent["gbs"] = ALG_BYTES[name] * B * n / (avg_ms * 1e-3) / 1e9
        kern[name] = ent
    dom = max((k for k in kern if k in ALG_BYTES), key=lambda k: kern[k]["share"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)      # DRAM bytes per launch from the committed ncu --set full capture
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["gbs"] / peak, "traffic": traffic,
                "traffic_source": "profiles/traffic.json <- ncu --set full capture of this build (profiles/r02_ncu_full.md)",
                "peak_source": peak_src,
                "share_of_step": kern[dom]["share"],
                "algorithmic_bytes_per_launch": ALG_BYTES[dom] * B * n,
                "step_frac_fwdbwd": (value / world) * (ALG_BYTES_FWDBWD if args.adjoint == "recompute"
                                                        else ALG_BYTES_TAPED) / 1e9 / peak,
                "step_frac_fwd": (fwd_value / world) * ALG_BYTES_FWD / 1e9 / peak}

    # ---------------- end-to-end with host buffers (pinned), copies inside the timed region
    # ONE flat pinned buffer and ONE cudaMemcpyAsync per direction and step (round 1 issued sixteen copies per step
    # and rank, and the 8-GPU e2e number collapsed to half the device rate).
    names = ("x", "v", "C", "F", "J")
    in_shapes = [tuple(getattr(state, k).shape) for k in names] + [tuple(action.shape)]
    out_shapes = [tuple(getattr(state, k).shape) for k in names] + \
                 [tuple(getattr(state, k).shape) for k in ("x", "v", "C", "F")] + [tuple(action.shape)]

    def numel(sh):
        r = 1
        for d in sh:
            r *= d
        return r

    def views(flat, shapes):
        out, o = [], 0
        for sh in shapes:
            out.append(flat[o:o + numel(sh)].view(sh))
            o += numel(sh)
        return out
    n_in, n_out = sum(numel(sh) for sh in in_shapes), sum(numel(sh) for sh in out_shapes)
    host_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
    host_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
    for v_, src in zip(views(host_in, in_shapes), [getattr(state, k) for k in names] + [action]):
        v_.copy_(src.detach().cpu())
    h2d, d2h = 4 * n_in, 4 * n_out

    # Three streams, double-buffered device inputs: the H2D of step i+1 and the D2H of step i-1 overlap the kernels
    # of step i (all copies stay inside the timed region; PCIe is full duplex).
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    s_comp = torch.cuda.current_stream(dev)
    dev_in = [torch.empty(n_in, dtype=torch.float32, device=dev) for _ in range(2)]
    dev_out = [torch.empty(n_out, dtype=torch.float32, device=dev) for _ in range(2)]
    ev_h2d = [torch.cuda.Event() for _ in range(2)]
    ev_comp = [torch.cuda.Event() for _ in range(2)]
    ev_d2h = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(i):
        slot = i % 2
        with torch.cuda.stream(s_h2d):
            s_h2d.wait_event(ev_comp[slot])          # the step that last read this slot has finished
            dev_in[slot].copy_(host_in, non_blocking=True)
            ev_h2d[slot].record(s_h2d)
        s_comp.wait_event(ev_h2d[slot])
        s_comp.wait_event(ev_d2h[slot])              # the slot's previous outputs have left the device
        vin = views(dev_in[slot], in_shapes)
        o, gr, _ = fwd_bwd(sim, state._replace(**dict(zip(names, vin[:5]))), vin[5], cot)
        torch.cat([getattr(o, k).detach().reshape(-1) for k in names] + [t.reshape(-1) for t in gr], out=dev_out[slot])
        ev_comp[slot].record(s_comp)
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_comp[slot])
            host_out.copy_(dev_out[slot], non_blocking=True)
            ev_d2h[slot].record(s_d2h)

    for i in range(0 if args.no_e2e else 2):
        e2e_step(i)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    g0.record()
    for i in range(1 if args.no_e2e else args.steps):
        e2e_step(i)
    s_comp.wait_stream(s_d2h)
    g1.record()
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    e2e_ms = max(g0.elapsed_time(g1), wall_ms)
    if dist is not None:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * units_per_step * args.steps / (e2e_ms * 1e-3)
    del dev_in, dev_out

    # ---------------- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, n_cpu, secs = cpu_oracle_rate(conf, 1, args.density, 2, 2, threads)
        cpu = {"value": rate, "unit": "particle-substeps/s", "cores": threads, "kind": "port",
               "sample": f"1 env x {n_cpu} particles x 2 substeps fwd+bwd (best of 2, {secs:.1f} s each), "
                         "torch-CPU oracle restating the reference (JAX not installable in this image)"}

    if rank == 0:
        line = {
            "metric": "particle-substeps/s fwd+bwd", "value": value, "unit": "particle-substeps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name, "state": f"every step starts from the scene after {args.settle} settle steps", "envs_per_gpu": B, "particles_per_env": n, "substeps": S,
                       "l2": f"inputs larger than L2 (state+checkpoints {B * n * 96 * (S + 1) / 1e9:.2f} GB per step)",
                       "adjoint": f"{args.adjoint}: " + ("substep residuals kept in HBM by the forward ("
                                   f"{L.ud_mpm_tape_bytes(C.byref(sim.params())) / 1e9:.2f} GB per step in flight)"
                                   if args.adjoint != "recompute" else
                                   "step input kept, the S substeps recomputed in the backward"),
                       "p2g_mode": "atomic" if args.p2g_mode == 0 else "deterministic",
                       "ckpt_window": sim.ckpt_window, "env_groups": sim.env_groups,
                       "collective": "none in the step (envs are independent); APG's policy-gradient all-reduce is "
                                     "timed beside it as `apg_update` when N > 1"},
            "forward_only": {"value": fwd_value, "unit": "particle-substeps/s"},
            "taped": taped,
            "e2e": {"value": e2e_value, "unit": "particle-substeps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "copies_per_step": 2, "host_gbs_per_rank": (h2d + d2h) / (e2e_ms / args.steps * 1e-3) / 1e9,
                    "host_gbs_all_ranks": world * (h2d + d2h) / (e2e_ms / args.steps * 1e-3) / 1e9, "numa": numa},
            "apg_update": None,
            "peak_hbm_bytes": peak_hbm,
            "gpu_launches": launches,
            "roofline": roofline,
            "kernels": {k: {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                        for k, v in kern.items()},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
    else:
        line = None

    # ---------------- the path's one collective, timed in the same run (N > 1), AFTER every other number is in `line`:
    # a watchdog prints the line without it if the leg (symmetric-memory rendezvous, flag barrier in peer memory) does
    # not come back, so that a multi-GPU run can never end without its JSON line
    emitted = threading.Event()

    def emit(apg):
        if not emitted.is_set():
            emitted.set()
            if line is not None:
                line["apg_update"] = apg
                print(json.dumps(line), flush=True)

    if dist is not None:
        finished = threading.Event()

        def watchdog():
            if not finished.wait(180):
                emit({"error": "apg_update leg did not finish within 180 s"})
                os._exit(0)
        threading.Thread(target=watchdog, daemon=True).start()
        try:
            apg_leg = apg_update_leg(dev, dist, world)
        except Exception as e:      # noqa: BLE001
            apg_leg = {"error": repr(e)[:200]}
        emit(apg_leg)
        threading.Timer(60, lambda: os._exit(0)).start()      # (a process group that cannot be torn down)
        try:
            dist.destroy_process_group()
        except Exception:           # noqa: BLE001
            pass
        finished.set()
        sys.stdout.flush()
        os._exit(0)
    emit(None)


if __name__ == "__main__":
    main()
