"""GPU: `unidom_b200.envs.ShapeElastoPlasticEnv.step_diff` (BASELINE configs[1]'s env: focus shift, 20 sub-actions x 16
substeps through the kernels, l2 reward) against the reference's own env run under oracle/jaxshim
(oracle/gen_golden.py::mpm_env_case -> tests/golden/ref_mpmenv_push.npz)."""
import os

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


def test_push_env_step_and_action_gradient_vs_reference(built_lib):
    from unidom_b200 import confs, envs
    d = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in np.load(os.path.join(util.GOLD, "ref_mpmenv_push.npz")).items()}
    B = d["in_x"].shape[0]
    conf = confs.shape_elasto_plastic_conf()
    env = envs.ShapeElastoPlasticEnv(conf, B, density=float(d["density"]), goal=d["goal"].numpy(), aux_reward=True)
    st = env.state
    assert util.rel_err(st.x, d["in_x"]) == 0.0                       # same add_box lattice, bit for bit
    a = d["actions"].to(env.device).requires_grad_(True)
    obs, reward, done, info = env.step_diff(a, st)
    ns = info["state"]
    (ga,) = torch.autograd.grad(reward.sum(), [a])
    # Bars: north_star's (state rtol 1e-4, gradients rtol 1e-3 / cosine 0.999), or K x the fp32 noise of the unmodified
    # reference on this very rollout (fp32 fixture vs the same run in fp64, util.env_floor) where that is higher.
    fl = util.env_floor("push")
    for k, ref in (("x", d["out_x"]), ("v", d["out_v"]), ("F", d["out_F"])):
        e, bar = util.rel_err(getattr(ns, k), ref), util.floor_bar(1e-4, fl("out_" + k))
        print(f"mpm env push {k}: cuda-vs-reference rel {e:.3e}  reference fp32-vs-fp64 floor {fl('out_' + k):.1e}  bar {bar:.1e}")
        assert e < bar, (k, e, bar)
    assert util.rel_err(ns.primitives[0].position, d["out_prim_pos"]) < 1e-5
    er = util.rel_err(reward, d["reward"])
    print(f"mpm env push reward {reward.tolist()} ref {d['reward'].tolist()} rel {er:.3e}")
    assert er < 1e-4
    # Per env.  The reference's norm_grad scrubs cotangents with nan_to_num, which maps +-inf to FLT_MAX; the global
    # norm then overflows and g / norm zeroes the WHOLE cotangent of that env for that step.  An inf arises when a
    # particle's weight product underflows (grad of p/m at a cell with m^2 == 0), and whether the numerator is 0 (NaN,
    # harmless) or denormal (inf) depends on rounding: the shim run itself flips between the two from run to run.  An
    # env whose reference gradient is exactly the direct contact term was such an instance and is reported, not compared.
    x = d["in_x"]
    a0 = d["actions"].clone().requires_grad_(True)
    contact = torch.sqrt(((a0[:, None, :3] - x) ** 2).sum(-1)).min(-1).values
    (gc,) = torch.autograd.grad((np.e ** (-contact)).sum(), [a0])
    compared = 0
    for b in range(B):
        ref = d["g_actions"][b]
        if float((ref - gc[b]).abs().max()) < 1e-6 * float(ref.abs().max()):
            print(f"env {b}: reference simulation gradient was zeroed by the inf -> FLT_MAX -> norm = inf scrub; ours {ga[b].tolist()}")
            continue
        cs, eg = util.cosine(ga[b], ref), util.rel_err(ga[b], ref)
        print(f"env {b}: action gradient cos {cs:.6f} rel {eg:.3e}")
        assert cs >= 0.999 and eg < 1e-3, (b, cs, eg)                      # north_star bar
        compared += 1
    assert compared >= 1
    eo = util.rel_err(obs, d["obs"])                                      # obs carries the velocities
    print(f"mpm env push obs: rel {eo:.3e}  floor {fl('obs'):.1e}")
    assert eo < util.floor_bar(1e-4, fl("obs"))


def test_whip_rope_env_two_steps_vs_reference(built_lib):
    """The reference's WhipRopeEnv (envs/whip_rope_env.py) at its shipped size under oracle/jaxshim
    (gen_golden.py::task_env_case): position control, focus shift, one 70-substep sub-action per env step, two env
    steps; rewards, states and the gradient of the summed rewards w.r.t. both actions."""
    from unidom_b200 import confs, envs
    d = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in np.load(os.path.join(util.GOLD, "ref_mpmenv_whip.npz")).items()}
    B = d["in_x"].shape[0]
    conf = confs.whip_rope_conf()
    env = envs.WhipRopeEnv(conf, B, goal=d["goal"].numpy())
    st = env.state
    assert st.x.shape[1] == d["in_x"].shape[1] == 67
    assert torch.equal(env.simulator.material.cpu(), d["material"]) and torch.equal(env.simulator.h.cpu(), d["h"])
    # the reference's reset applies a threefry-drawn xz shift: take the shifted scene from the fixture
    p = st.primitives[0]._replace(position=d["in_prim0_pos"].to(env.device), rotation=d["in_prim0_rot"].to(env.device))
    base = (d["in_x"] - d["in_x"].mean(1, keepdim=True)) - (st.x.cpu() - st.x.cpu().mean(1, keepdim=True))
    assert float(base.abs().max()) < 1e-6                               # same rope lattice up to the shift
    st = st._replace(x=d["in_x"].to(env.device), primitives=[p])
    a = d["actions"].to(env.device).requires_grad_(True)
    fl = util.env_floor("whip")      # fp32 noise of the unmodified reference on this rollout (fp32 fixture vs fp64 run)
    total, s = 0, st
    for t in range(2):
        obs, reward, done, info = env.step_diff(a[t], s)
        s = info["state"]
        total = total + reward.sum()
        ex, ev = util.rel_err(s.x, d[f"x{t}"]), util.rel_err(s.v, d[f"v{t}"])
        er = util.rel_err(reward, d[f"reward{t}"])
        print(f"whip env step {t}: x rel {ex:.3e} (floor {fl(f'x{t}'):.1e}) v rel {ev:.3e} (floor {fl(f'v{t}'):.1e}) "
              f"reward {reward.tolist()} ref {d[f'reward{t}'].tolist()} rel {er:.3e}")
        assert ex < util.floor_bar(1e-4, fl(f"x{t}")) and ev < util.floor_bar(1e-4, fl(f"v{t}")) and er < 1e-4
        assert util.rel_err(s.primitives[0].position, d[f"prim0_pos{t}"]) < 1e-5
    (ga,) = torch.autograd.grad(total, [a])
    cs, eg = util.cosine(ga, d["g_actions"]), util.rel_err(ga, d["g_actions"])
    print(f"whip env action gradient cos {cs:.6f} rel {eg:.3e}  floor {fl('g_actions'):.1e}")
    assert cs >= 0.999 and eg < 1e-3
    assert util.rel_err(obs, d["obs"]) < util.floor_bar(1e-4, fl("obs"))


def test_pour_water_env_two_steps_vs_reference(built_lib):
    """The reference's PourWaterEnv (envs/pour_water_env.py) at its shipped size (702 liquid particles, two bowl colliders
    with the container SDF and finite-difference normals) under oracle/jaxshim: two env steps of 23 substeps."""
    from unidom_b200 import confs, envs
    d = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in np.load(os.path.join(util.GOLD, "ref_mpmenv_pour.npz")).items()}
    B, n = d["in_x"].shape[:2]
    conf = confs.pour_water_conf()
    env = envs.PourWaterEnv(conf, B, goal=d["goal"].numpy(), points=d["in_x"][0].numpy())
    st = env.state
    assert torch.equal(env.simulator.material.cpu(), d["material"]) and torch.equal(env.simulator.h.cpu(), d["h"])
    prims = [p._replace(position=d[f"in_prim{q}_pos"].to(env.device), rotation=d[f"in_prim{q}_rot"].to(env.device))
             for q, p in enumerate(st.primitives)]
    st = st._replace(x=d["in_x"].to(env.device), primitives=prims)
    a = d["actions"].to(env.device).requires_grad_(True)
    fl = util.env_floor("pour")      # fp32 noise of the unmodified reference on this rollout (fp32 fixture vs fp64 run)
    total, s = 0, st
    for t in range(2):
        obs, reward, done, info = env.step_diff(a[t], s)
        s = info["state"]
        total = total + reward.sum()
        ex, ev = util.rel_err(s.x, d[f"x{t}"]), util.rel_err(s.v, d[f"v{t}"])
        er = util.rel_err(reward, d[f"reward{t}"])
        print(f"pour env step {t}: x rel {ex:.3e} (floor {fl(f'x{t}'):.1e}) v rel {ev:.3e} (floor {fl(f'v{t}'):.1e}) "
              f"reward {reward.tolist()} ref {d[f'reward{t}'].tolist()} rel {er:.3e}")
        # v: the bowl's finite-difference normals (d = 1e-6 in fp32, primitives.py:129-143) quantise to a few ulps of
        # the SDF, so particles touching the bowl carry that noise in v -- in the reference's own fp32 run as well:
        # the bar is K x the measured fp32-vs-fp64 gap of the reference
        assert ex < util.floor_bar(1e-4, fl(f"x{t}")) and ev < util.floor_bar(1e-4, fl(f"v{t}")) and er < 1e-4
        for q in range(2):
            assert util.rel_err(s.primitives[q].position, d[f"prim{q}_pos{t}"]) < 1e-5
            assert util.rel_err(s.primitives[q].rotation, d[f"prim{q}_rot{t}"]) < 1e-5
    (ga,) = torch.autograd.grad(total, [a])
    cs, eg = util.cosine(ga, d["g_actions"]), util.rel_err(ga, d["g_actions"])
    print(f"pour env action gradient cos {cs:.6f} rel {eg:.3e}  max|ref| {float(d['g_actions'].abs().max()):.3e}")
    # the action gradient reaches the liquid only through the bowl collider, whose normals are fp32 central differences
    # with d = 1e-6 (a few ulps of the SDF): two fp32 evaluations of the same formulas differ at the printed floor
    print(f"pour env action gradient: reference fp32-vs-fp64 floor {fl('g_actions'):.1e}  bar {util.floor_bar(1e-3, fl('g_actions')):.1e}")
    assert cs >= 0.999 and eg < util.floor_bar(1e-3, fl("g_actions"))
    assert util.rel_err(obs, d["obs"]) < util.floor_bar(1e-4, fl("obs"))


def test_shape_rope_env_step_vs_reference(built_lib):
    """The reference's ShapeRopeEnv (envs/shape_rope_env.py) at its shipped size (582 plastic particles, n_grid 128,
    30 sub-actions x 133 substeps = 3 990 substeps per env step) from the state its reset() leaves after the two random
    pushes: one env step, reward and action gradient."""
    path = os.path.join(util.GOLD, "ref_mpmenv_rope.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    from unidom_b200 import confs, envs
    d = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in np.load(path).items()}
    B = d["in_x"].shape[0]
    conf = confs.shape_rope_conf()
    env = envs.ShapeRopeEnv(conf, B, goal=d["goal"].numpy())
    st = env.state
    assert st.x.shape[1] == d["in_x"].shape[1] == 582
    assert torch.equal(env.simulator.material.cpu(), d["material"]) and torch.equal(env.simulator.h.cpu(), d["h"])
    p = st.primitives[0]._replace(position=d["in_prim0_pos"].to(env.device), rotation=d["in_prim0_rot"].to(env.device))
    st = st._replace(primitives=[p], **{k: d[f"in_{k}"].to(env.device) for k in ("x", "v", "C", "F", "J")})
    a = d["actions"].to(env.device).requires_grad_(True)
    obs, reward, done, info = env.step_diff(a[0], st)
    s = info["state"]
    (ga,) = torch.autograd.grad(reward.sum(), [a])
    fl = util.env_floor("rope")      # fp32 noise of the unmodified reference over these 3 990 substeps (fp32 fixture vs fp64 run)
    ex, er = util.rel_err(s.x, d["x0"]), util.rel_err(reward, d["reward0"])
    print(f"rope env: x rel {ex:.3e} (floor {fl('x0'):.1e}) reward {reward.tolist()} ref {d['reward0'].tolist()} rel {er:.3e} "
          f"(floor {fl('reward0'):.1e})")
    assert ex < util.floor_bar(1e-4, fl("x0")) and er < util.floor_bar(1e-4, fl("reward0"))
    assert util.rel_err(s.primitives[0].position, d["prim0_pos0"]) < 1e-5
    cs, eg = util.cosine(ga, d["g_actions"]), util.rel_err(ga, d["g_actions"])
    print(f"rope env action gradient cos {cs:.6f} rel {eg:.3e} (floor {fl('g_actions'):.1e})  ours {ga.tolist()} ref {d['g_actions'].tolist()}")
    assert cs >= 0.999 and eg < util.floor_bar(1e-3, fl("g_actions"))
    # the reference's own reset continues with random pushes: they run and keep the rope on the table
    env.state = s
    s2 = env.random_push(step=1, rng=np.random.RandomState(0))
    assert torch.isfinite(s2.x).all() and float(s2.x[..., 1].min()) > -1e-3
