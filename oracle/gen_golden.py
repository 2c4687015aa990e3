#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference sources.

    python oracle/gen_golden.py            # in the build container (reference mounted at /root/reference)

The reference (DaXBench) is Python + JAX; jax/jaxlib are not installable in this image.  This script puts
`oracle/jaxshim` (a torch-backed stand-in for the jax / optax APIs the path uses) in front of sys.path and
imports /root/reference/DaXBench/daxbench/core/engine/{mpm_simulator,cloth_simulator,svd_safe_batch}.py and
primitives/{primitives,box,container}.py AS THEY ARE, then calls the reference's own
`SimpleMPMSimulator.step_jax` / `ClothSimulator.step_jax` and `jax.grad` through the reference's own
custom_vjp rules on small seeded scenes.  With a real JAX install (`--real-jax`) the same script runs
against jax itself.

TEST INFRASTRUCTURE ONLY.  The fixtures (inputs, outputs, cotangents, gradients, conf scalars) travel with
the repo; /root/reference is never read at test time.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("UNIDOM_REFERENCE", "/root/reference/DaXBench")

STATE_F = ("x", "v", "C", "F", "J", "friction", "mu", "lamda")
PRIM_F = ("size", "friction", "softness", "position", "rotation", "v", "w", "action_buffer", "action_scale")
CLOTH_F = ("x", "v", "primitive0", "primitive1", "action0", "action1", "stiffness", "mu")


def load_reference(real_jax):
    if not real_jax:
        sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    sys.path.insert(0, REF)
    import jax  # noqa: F401
    from daxbench.core.engine import cloth_simulator, mpm_simulator
    from daxbench.core.engine.primitives import box, container, primitives
    return mpm_simulator, cloth_simulator, primitives, box, container


def npy(a):
    return np.asarray(a)


# --------------------------------------------------------------------------------------------- MPM
class MPMConf:
    """Scalar fields of the reference's per-task DefaultConf (envs/shape_elasto_plastic.py:23-54 etc.)."""

    def __init__(self, jnp, key, n_grid, res, dt, steps, E, nu, ground_friction, n_primitive, gravity=(0, -9.8, 0)):
        self.seed = 1
        self.key = key
        self.n_primitive = n_primitive
        self.ground_friction = ground_friction
        self.n_grid, self.steps, self.dt = n_grid, steps, dt
        self.E, self.nu = E, nu
        self.res = tuple(res)
        self.dx, self.inv_dx = 1 / n_grid, float(n_grid)
        self.p_vol, self.p_rho = (self.dx * 0.5) ** 2, 1
        self.p_mass = self.p_vol * self.p_rho
        self.gravity = jnp.array(list(gravity))


def mpm_case(mods, name, material, n_prim, pos_control, sdf_kind, steps, seed, cot_scale, ylow=0.02, v_scale=0.3,
             c_scale=2.0, f_scale=0.05, B=2, grads=True, zero_rot=False):
    import jax
    import jax.numpy as jnp
    mpm, _, prim, box, container = mods
    rng = np.random.RandomState(seed)
    conf = MPMConf(jnp, jax.random.PRNGKey(0), n_grid=96, res=(48, 32, 48), dt=2e-4, steps=steps, E=2, nu=0.2,
                   ground_friction=2, n_primitive=n_prim)
    prim.set_sdf(box._sdf_batch if sdf_kind == 0 else container._sdf_batch)
    sim = mpm.SimpleMPMSimulator(conf, B, use_position_control=pos_control)
    sim.key_global = jax.random.PRNGKey(1)
    # scene: the reference's own add_box lattice (material != 0) or seeded points (the liquid branch draws
    # from jax.random, which the shim does not reproduce)
    if material == 0:
        size = np.array([0.1, 0.06, 0.08], np.float32)
        n_pts = int(np.prod(size.astype(np.float64)) * conf.n_grid ** 3)
        pts = (rng.uniform(size=(n_pts, 3)).astype(np.float32) * 2 - 1) * (np.float32(0.5) * size) \
            + np.array([0.25, ylow, 0.25], np.float32)
        state = sim.add_box_from_points(conf, None, jnp.array(pts), hardness=1.0, material=0)
    else:
        state = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.1, 0.06, 0.08], init_pos=[0.25, ylow, 0.25],
                            z_rotation_angle=0, material=material, density=1)
    prims = []
    if sdf_kind == 0:
        specs = [(0.9, [0.015, 0.06, 0.015], [0.25, 0.01, 0.205]), (0.5, [0.02, 0.03, 0.01], [0.29, 0.02, 0.27])]
    else:  # bowls: size = (r, h, t)
        specs = [(0.9, [0.06, 0.02, 0.004], [0.25, 0.06, 0.25]), (0.5, [0.04, 0.01, 0.004], [0.29, 0.05, 0.27])]
    for q in range(n_prim):
        fr, size, pos = specs[q]
        p = prim.create_primitive(conf, friction=fr, softness=666, color=[0.5, 0.5, 0.5], size=size, init_pos=pos)
        prims.append(p._replace(action_scale=p.action_scale * 0.012))
    state = state._replace(primitives=prims)
    state = sim.reset_jax(state)
    n = state.x.shape[1]
    f32 = np.float32
    x = npy(state.x) + (rng.randn(B, n, 3) * 1e-3).astype(f32)
    v = (rng.randn(B, n, 3) * v_scale).astype(f32)
    Cm = (rng.randn(B, n, 3, 3) * c_scale).astype(f32)
    F = (np.eye(3, dtype=f32)[None, None] + f_scale * rng.randn(B, n, 3, 3)).astype(f32)
    state = state._replace(x=jnp.array(x), v=jnp.array(v), C=jnp.array(Cm), F=jnp.array(F))
    action = rng.rand(B, 6 * n_prim).astype(f32) * 2.4 - 1.2          # some entries beyond the [-1, 1] clip
    action.reshape(B, n_prim, 6)[:, :, 3:] *= 0.0 if zero_rot else 0.3
    action = jnp.array(action)

    out = {"material": npy(sim.material).astype(np.int32), "h": npy(sim.h).astype(np.float32), "action": npy(action),
           "conf": np.array([conf.n_grid, *conf.res, conf.steps, n_prim, int(pos_control), sdf_kind], np.int64),
           "conf_f": np.array([conf.dt, conf.E, conf.nu, conf.ground_friction, *npy(conf.gravity)], np.float64)}
    for k in STATE_F:
        out["in_" + k] = npy(getattr(state, k))
    for q in range(n_prim):
        for k in PRIM_F:
            out[f"in_p{q}_{k}"] = npy(getattr(state.primitives[q], k)).astype(np.float32)

    new_state, _ = sim.step_jax(state, action)
    for k in STATE_F:
        out["out_" + k] = npy(getattr(new_state, k))
    for q in range(n_prim):
        for k in PRIM_F:
            out[f"out_p{q}_{k}"] = npy(getattr(new_state.primitives[q], k)).astype(np.float32)

    if grads:
        S = steps
        cot = {"x": rng.randn(B, n, 3), "v": rng.randn(B, n, 3) * 0.1, "C": rng.randn(B, n, 3, 3) * 1e-3,
               "F": rng.randn(B, n, 3, 3) * 0.1}
        for q in range(n_prim):
            cot[f"p{q}_position"] = rng.randn(B, S, 3)
            cot[f"p{q}_rotation"] = rng.randn(B, S, 4)
        cot = {k: (a * cot_scale).astype(f32) for k, a in cot.items()}
        for k, a in cot.items():
            out["cot_" + k] = a

        def loss(inp):
            st, act = inp
            ns, _ = sim.step_jax(st, act)
            L = 0.0
            for k in ("x", "v", "C", "F"):
                L = L + (getattr(ns, k) * jnp.array(cot[k])).sum()
            for q in range(n_prim):
                L = L + (ns.primitives[q].position * jnp.array(cot[f"p{q}_position"])).sum()
                L = L + (ns.primitives[q].rotation * jnp.array(cot[f"p{q}_rotation"])).sum()
            return L

        g_state, g_action = jax.grad(loss, allow_int=True)((state, action))
        out["g_action"] = npy(g_action)
        for k in ("x", "v", "C", "F", "friction", "mu", "lamda"):
            out["g_" + k] = npy(getattr(g_state, k))
        for q in range(n_prim):
            for k in ("size", "friction", "position", "rotation", "action_scale"):
                out[f"g_p{q}_{k}"] = npy(getattr(g_state.primitives[q], k)).astype(np.float32)
    path = os.path.join(GOLD, f"ref_mpm_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: n={n} B={B} S={steps}  max|out_x-in_x|={np.abs(out['out_x'] - out['in_x']).max():.3e}"
          + (f"  |g_x|max={np.abs(out['g_x']).max():.3e} |g_action|max={np.abs(out['g_action']).max():.3e}" if grads else ""))


# ------------------------------------------------------------------------------------------- cloth
class ClothConf:
    """envs/fold_cloth3_env.py:18-37."""
    N = 80
    cell_size = 1.0 / N
    gravity = 0.5
    stiffness = 900
    damping = 2
    dt = 2e-3
    max_v = 2.
    small_num = 1e-8
    mu = 0.5
    seed = 1
    size = int(N / 5.0)
    mem_saving_level = 2


def fold_cloth_mask(conf):
    """envs/fold_cloth3_env.py:51-56."""
    import jax.numpy as jnp
    N, size = conf.N, conf.size
    cloth_mask = jnp.zeros((N, N))
    cloth_mask = cloth_mask.at[size * 2:size * 3, size * 2:size * 4].set(1)
    return cloth_mask


def _closure_vars(fn):
    return dict(zip(fn.__code__.co_freevars, [c.cell_contents for c in fn.__closure__]))


def cloth_inner_step(sim):
    """The reference's per-substep `step_wrapper` (cloth_simulator.py:228-255), fished out of the closures of
    `sim.step_jax` (vmap -> robot_step_wrapper -> robot_step) without touching the source: robot_step hard-codes
    50 substeps (:176), and 50 chaotic substeps are too long a window for tight parity."""
    f = _closure_vars(sim.step_jax)["f"]                   # the function vmap wraps
    robot_step = _closure_vars(f.fun)["robot_step"]
    return _closure_vars(robot_step)["step_wrapper"]


def cloth_case(mods, name, float_stiffness, seed, from_reset, B=2, n_calls=1, grads=True, window=0, contact=False):
    import jax
    import jax.numpy as jnp
    _, cl, *_ = mods
    rng = np.random.RandomState(seed)
    conf = ClothConf()
    mask = fold_cloth_mask(conf)
    sim = cl.ClothSimulator(conf, B, lambda x, v, i, j: v, mask)      # identity collision (cloth_env.py:239-243)
    st = sim.reset_jax()
    f32 = np.float32
    P = st.x.shape[1]
    if not from_reset:
        x = npy(st.x) + 0.002 * rng.randn(B, P, 3).astype(f32)
        lift = (np.abs(x[..., 1]) * 3 + 0.01 * rng.rand(B, P)) * (rng.rand(B, P) > 0.5)
        x[..., 1] = lift
        v = 0.05 * rng.randn(B, P, 3).astype(f32)
        p0 = np.concatenate([x[:, 100], np.full((B, 1), 0.02)], axis=1).astype(f32)
        p1 = np.concatenate([x[:, 300] + 0.004, np.full((B, 1), 0.015)], axis=1).astype(f32)
        st = st._replace(x=jnp.array(x.astype(f32)), v=jnp.array(v), primitive0=jnp.array(p0), primitive1=jnp.array(p1))
    if float_stiffness:
        st = st._replace(stiffness=jnp.array((900.0 + 300 * rng.rand(B)).astype(f32)))
    st = st._replace(mu=jnp.array((0.3 + 0.4 * rng.rand(B)).astype(f32)))
    acts = np.array([[0.3, 0.5, -0.2, 0.0, -0.1, 0.2, 0.4, 0.3], [2.6, -0.4, 0.1, 1.0, 0.0, 0.0, 0.0, 0.0],
                     [0.0, 0.06, 0.0, 0.2, 0.5, 0.1, -3.0, 0.0]], f32)[:B]
    if from_reset:   # grab a corner node and lift it (a pick-and-place approach phase, suction 1 then 0)
        x0 = npy(st.x)
        p0 = np.concatenate([x0[:, 0], np.full((B, 1), 0.01)], axis=1).astype(f32)
        st = st._replace(primitive0=jnp.array(p0))
        acts = np.array([[0.1, 0.6, 0.1, 0.0, 0, 0, 0, 1.0], [0.0, 0.8, -0.2, 0.0, 0, 0, 0, 1.0]], f32)[:B]
    if contact:     # the oracle's state 44 substeps into a violent sub-action (ground contacts, |v| = max_v, closed gripper)
        import torch
        cs = torch.load(os.path.join(GOLD, "cloth_contact_state.pt"))
        B = 1
        st = st._replace(**{k: jnp.array(cs[k][None].numpy()) for k in ("x", "v", "primitive0", "primitive1", "mu")})
        st = jax.tree_util.tree_map(lambda l: l[:1], st)
        if float_stiffness:
            st = st._replace(stiffness=jnp.array(cs["stiffness"][None].numpy().astype(f32)))
        acts = np.concatenate([cs["action0"][:3].numpy() * 50, cs["action0"][3:].numpy(),
                               cs["action1"][:3].numpy() * 50, cs["action1"][3:].numpy()]).astype(f32)[None]
    action = jnp.array(acts)
    out = {"action": npy(action), "mask": npy(mask).astype(np.int32), "n_calls": np.array(n_calls),
           "window": np.array(window), "stiffness_is_float": np.array(int(float_stiffness))}
    for k in CLOTH_F:
        out["in_" + k] = npy(getattr(st, k))

    inner = cloth_inner_step(sim) if window else None

    def run_window_env(s, action):
        # robot_step's prologue (cloth_simulator.py:168-169), then `window` substeps of the reference's step_wrapper
        action0 = action.at[:3].set(action[:3].clip(-2, 2) / 50.)[:4]
        action1 = action.at[4:7].set(action[4:7].clip(-2, 2) / 50.)[4:8]
        s = s._replace(action0=action0, action1=action1)
        for i in range(window):
            s = inner(i, s)
        return s

    def run(s, a):
        if window:
            return jax.vmap(run_window_env)(s, a)
        for _ in range(n_calls):
            s, _ = sim.step_jax(s, a)
        return s

    ns = run(st, action)
    for k in CLOTH_F:
        out["out_" + k] = npy(getattr(ns, k))
    if grads:
        cot = {"x": rng.randn(B, P, 3).astype(f32), "v": rng.randn(B, P, 3).astype(f32),
               "primitive0": rng.randn(B, 4).astype(f32), "primitive1": rng.randn(B, 4).astype(f32)}
        for k, a in cot.items():
            out["cot_" + k] = a

        def loss(inp):
            s, a = inp
            o = run(s, a)
            return sum((getattr(o, k) * jnp.array(cot[k])).sum() for k in cot)

        gs, ga = jax.grad(loss, allow_int=True)((st, action))
        out["g_action"] = npy(ga)
        for k in ("x", "v", "primitive0", "primitive1", "mu") + (("stiffness",) if float_stiffness else ()):
            out["g_" + k] = npy(getattr(gs, k))
    path = os.path.join(GOLD, f"ref_cloth_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: P={P} B={B} calls={n_calls}  max|dx|={np.abs(out['out_x'] - out['in_x']).max():.3e}"
          + (f"  |g_x|max={np.abs(out['g_x']).max():.3e} |g_action|max={np.abs(out['g_action']).max():.3e}" if grads else ""))


def cloth_env_case(name, ep_len, B, seed):
    """Env level (the callers of the step): the reference's FoldCloth3Env.step_diff (cloth_env.py:204-231) with
    get_pnp_actions / calc_chamfer / get_obs as they are, driven by the APG rollout of apg.py:177-215 with the
    policy MLP written in torch (brax/flax are not importable; weights and sampling noise are fixtures)."""
    sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    import stubs
    stubs.install()
    import torch
    import jax
    import jax.numpy as jnp
    from daxbench.core.envs import fold_cloth3_env
    from daxbench.core.utils import util as rutil
    sys.path.insert(0, ROOT)
    from unidom_b200 import apg
    rng = np.random.RandomState(seed)
    env = fold_cloth3_env.FoldCloth3Env(batch_size=B, aux_reward=True)
    obs0, st = env.reset(jax.random.PRNGKey(0))
    shift = rng.randn(2).astype(np.float32) * 0.05                        # cloth_env.py:183 with a seeded numpy draw
    x0 = np.asarray(env.simulator.reset_jax().x).copy()
    x0[..., 0] += shift[0]
    x0[..., 2] += shift[1]
    st = st._replace(x=jnp.array(x0))
    params = apg.init_policy(env.observation_size, env.action_size, seed=seed)
    params[-1] = params[-1] + torch.tensor([0.0, 0.0, 0.0, 0.6, 0.0, 0.4, -2, -2, -2, -2, -2, -2])   # aim at the cloth
    eps = torch.from_numpy(rng.randn(ep_len, B, env.action_size).astype(np.float32))
    out = {"goal": np.asarray(env.goal), "eps": eps.numpy(), "shift": shift, "ep_len": np.array(ep_len),
           "policy_seed": np.array(seed), "param5": params[-1].numpy()}      # weights are regenerated from the seed
    for k in CLOTH_F:
        out["in_" + k] = np.asarray(getattr(st, k))
    req = [p.clone().requires_grad_(True) for p in params]
    rewards, state = [], st
    for t in range(ep_len):
        obs = env.get_obs(state)
        actions = apg.sample_actions(apg.policy_apply(req, obs.t), eps[t], True)
        out[f"actions{t}"] = actions.detach().numpy()
        out[f"obs{t}"] = np.asarray(obs)
        if t == 0:
            out["pnp0"] = np.asarray(env.get_pnp_actions(jax.numpy.array(actions.detach().numpy()), state))
            out["chamfer0"] = np.asarray(rutil.calc_chamfer(state.x, env.goal))
        _, reward, done, info = env.step_diff(jax.Array(actions), state)
        state = info["state"]
        rewards.append(reward.t)
        out[f"reward{t}"] = reward.t.detach().numpy()
        out[f"real_reward{t}"] = np.asarray(info["real_reward"].t.detach())
        out[f"x{t + 1}"] = state.x.t.detach().numpy()
        out[f"v{t + 1}"] = state.v.t.detach().numpy()
        out[f"primitive0_{t + 1}"] = state.primitive0.t.detach().numpy()
    loss = -torch.stack(rewards).mean()
    grads = torch.autograd.grad(loss, req)
    out["loss"] = loss.detach().numpy()
    # the first two layers' gradients (3.7 MB) follow from the last layer's by the same torch MLP backward on both
    # sides; the fixture keeps layer 3 (W3, b3) and the full-gradient norm
    out["gparam4"], out["gparam5"] = grads[4].numpy(), grads[5].numpy()
    out["gnorm"] = np.array(float(torch.sqrt(sum((g * g).sum() for g in grads))))
    if ep_len > 1:
        # The reference's OWN sensitivity (the noise floor the GPU test asserts against): the same free-running
        # rollout of the unmodified reference from input positions perturbed by 1e-7 (about one fp32 ulp of x).
        prng = np.random.RandomState(seed + 1000)
        xp = (x0 + 1e-7 * prng.randn(*x0.shape)).astype(np.float32)
        req_p = [p.clone().requires_grad_(True) for p in params]
        rs, state = [], st._replace(x=jnp.array(xp))
        for t in range(ep_len):
            obs = env.get_obs(state)
            actions = apg.sample_actions(apg.policy_apply(req_p, obs.t), eps[t], True)
            _, reward, done, info = env.step_diff(jax.Array(actions), state)
            state = info["state"]
            rs.append(reward.t)
            out[f"pert_x{t + 1}"] = state.x.t.detach().numpy()
            out[f"pert_reward{t}"] = reward.t.detach().numpy()
        loss_p = -torch.stack(rs).mean()
        gp = torch.autograd.grad(loss_p, req_p)
        out["pert_loss"] = loss_p.detach().numpy()
        out["pert_gparam4"], out["pert_gparam5"] = gp[4].numpy(), gp[5].numpy()
        cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm()))
        print(f"reference vs reference with 1e-7 input perturbation: loss {float(loss):.6f} / {float(loss_p):.6f}, "
              f"policy-gradient cosine {cos(grads[4], gp[4]):.6f} / {cos(grads[5], gp[5]):.6f}")
    path = os.path.join(GOLD, f"ref_clothenv_{name}.npz")
    np.savez_compressed(path, **out)
    gn = float(torch.sqrt(sum((g * g).sum() for g in grads)))
    print(f"wrote {path}: ep_len={ep_len} B={B} loss={float(loss):.6f} |grad|={gn:.4e}")


def cloth_env_para_case(name, B, seed, it):
    """BASELINE configs[3] (GenDOM parameter-aware APG): the reference's FoldCloth1ParaEnv (fold_cloth1_para_env.py:39,
    cloth_env_para.py:98-135 get_obs with the normalised stiffness, :199-233 step_diff) with the stiffness drawn the
    way apg_para.py:326-329 draws it for iteration `it`, one APG rollout step (apg_para.py:201-240) with the torch
    policy; outputs obs, reward, the policy gradient and d loss / d stiffness."""
    sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    import stubs
    stubs.install()
    import torch
    import jax
    import jax.numpy as jnp
    from daxbench.core.envs import fold_cloth1_para_env
    sys.path.insert(0, ROOT)
    from unidom_b200 import apg
    rng = np.random.RandomState(seed)
    np.random.seed(it)                                                    # apg_para.py:326-329, verbatim
    stiffness = np.random.uniform(200, 1800)
    eval_mm = [100, 2000]
    env = fold_cloth1_para_env.FoldCloth1ParaEnv(batch_size=B, aux_reward=True, seed=0, stiffness=stiffness,
                                                 eval_min_max_stiff=eval_mm)
    _, st = env.reset(jax.random.PRNGKey(0))
    shift = rng.randn(2).astype(np.float32) * 0.05
    x0 = np.asarray(env.simulator.reset_jax().x).copy()
    x0[..., 0] += shift[0]
    x0[..., 2] += shift[1]
    stiff_t = torch.as_tensor(np.asarray(st.stiffness)).clone().requires_grad_(True)
    st = st._replace(x=jnp.array(x0), stiffness=jax.Array(stiff_t))
    params = apg.init_policy(env.observation_size, env.action_size, seed=seed)
    params[-1] = params[-1] + torch.tensor([0.0, 0.0, 0.0, 0.6, 0.0, 0.4, -2, -2, -2, -2, -2, -2])
    eps = torch.from_numpy(rng.randn(1, B, env.action_size).astype(np.float32))
    out = {"goal": np.asarray(env.goal), "eps": eps.numpy(), "shift": shift, "policy_seed": np.array(seed),
           "param5": params[-1].numpy(), "stiffness_draw": np.array(stiffness, np.float64), "it": np.array(it),
           "eval_min_max_stiff": np.array(eval_mm, np.float64)}
    for k in CLOTH_F:
        v = getattr(st, k)
        out["in_" + k] = (v.t.detach().numpy() if hasattr(v, "t") else np.asarray(v))
    req = [p.clone().requires_grad_(True) for p in params]
    eval_tile = np.tile(np.array(eval_mm), (B, 1))
    obs = env.get_obs(st, eval_min_max_stiff=eval_tile)                     # apg_para.py:205-208
    out["obs0"] = obs.t.detach().numpy()
    actions = apg.sample_actions(apg.policy_apply(req, obs.t), eps[0], True)
    out["actions0"] = actions.detach().numpy()
    obs1, reward, done, info = env.step_diff(jax.Array(actions), st)
    state = info["state"]
    out["obs1"] = obs1.t.detach().numpy()
    out["reward0"] = reward.t.detach().numpy()
    out["x1"] = state.x.t.detach().numpy()
    out["stiffness1"] = state.stiffness.t.detach().numpy()
    loss = -reward.t.mean()
    grads = torch.autograd.grad(loss, req + [stiff_t], allow_unused=True)
    out["loss"] = loss.detach().numpy()
    out["gparam4"], out["gparam5"] = grads[4].numpy(), grads[5].numpy()
    out["gnorm"] = np.array(float(torch.sqrt(sum((g * g).sum() for g in grads[:6]))))
    out["g_stiffness"] = np.zeros(B, np.float32) if grads[6] is None else grads[6].numpy()
    path = os.path.join(GOLD, f"ref_clothenv_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: B={B} stiffness={stiffness:.4f} loss={float(loss):.6f} |grad|={float(out['gnorm']):.4e} "
          f"g_stiffness={out['g_stiffness']}")


def cloth_unfold_case(name, B, seed, key_seed=1):
    """The unfold tasks' reset (envs/unfold_cloth1_env.py:56-82): lattice + threefry normal noise * 1e-4, then ONE
    pick-and-place fold between two nodes drawn from np.random (2 000 substeps through step_diff), run from the
    reference's own UnfoldCloth1Env under the shim with np.random seeded.  Also the same reset of the unmodified
    reference from lattice positions perturbed by 1e-7 (about one fp32 ulp): the chaotic rollout's own sensitivity,
    the floor the GPU test scales its bar with."""
    sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    import stubs
    stubs.install()
    import torch
    import jax
    import jax.numpy as jnp
    from jax import random
    from daxbench.core.envs import unfold_cloth1_env as mod
    env = mod.UnfoldCloth1Env(batch_size=B, seed=1)
    np.random.seed(seed)
    obs, st = env.reset(jax.random.PRNGKey(key_seed))
    # what reset() drew, recomputed the same way for the fixture
    init = env.simulator.reset_jax()
    k, _ = random.split(jax.random.PRNGKey(key_seed))
    noisy = init.x + random.normal(k, init.x.shape) * 0.0001
    np.random.seed(seed)
    st_point = np.random.randint(0, init.x.shape[1], size=(B,))
    ed_point = np.random.randint(0, init.x.shape[1], size=(B,))
    out = {"lattice_x": np.asarray(init.x), "noisy_x": np.asarray(noisy), "st_point": st_point, "ed_point": ed_point,
           "x": np.asarray(st.x), "v": np.asarray(st.v), "primitive0": np.asarray(st.primitive0), "obs": np.asarray(obs),
           "mu": np.asarray(st.mu), "seed": np.array(seed), "key_seed": np.array(key_seed), "goal": np.asarray(env.goal)}
    # sensitivity of the reference itself: the same fold from positions perturbed by 1e-7
    prng = np.random.RandomState(seed + 1000)
    xp = (np.asarray(noisy) + 1e-7 * prng.randn(*noisy.shape)).astype(np.float32)
    bidx = jnp.arange(B)
    s0 = init._replace(x=jnp.array(xp))
    actions = jnp.concatenate((s0.x[bidx, st_point], s0.x[bidx, ed_point]), axis=-1)
    _, _, _, info = env.step_diff(actions, s0)
    out["pert_x"] = np.asarray(info["state"].x)
    path = os.path.join(GOLD, f"ref_clothenv_{name}.npz")
    np.savez_compressed(path, **out)
    d = np.abs(out["pert_x"] - out["x"]).max()
    print(f"wrote {path}: B={B} fold {st_point}->{ed_point}, |x - lattice|max={np.abs(out['x'] - out['lattice_x']).max():.3f}, "
          f"reference vs 1e-7-perturbed reference |dx|max={d:.3e}")


F64 = False   # --f64: the env cases once more in float64 FROM THE FIXTURE'S INPUTS -> ref_mpmenv_<name>_f64.npz.  The gap
              # between the fp32 fixture and this run is the fp32 noise of the UNMODIFIED reference itself on that rollout:
              # the floor the env-level GPU tests print and scale their bars with (tests/util.py::env_floor).


def _f64_override(name, st, jnp):
    """fp64 run: start from exactly the state the fp32 fixture started from."""
    import torch
    d = np.load(os.path.join(GOLD, f"ref_mpmenv_{name}.npz"))
    rep = {k: jnp.array(torch.from_numpy(d["in_" + k]).double()) for k in ("x", "v", "C", "F", "J") if "in_" + k in d}
    prims = list(st.primitives)
    for q in range(len(prims)):
        if f"in_prim{q}_pos" in d:
            prims[q] = prims[q]._replace(position=jnp.array(torch.from_numpy(d[f"in_prim{q}_pos"]).double()),
                                         rotation=jnp.array(torch.from_numpy(d[f"in_prim{q}_rot"]).double()))
    return st._replace(primitives=prims, **rep), d["actions"].astype(np.float64)


def mpm_env_case(name, B, density, seed):
    """Env level for MPM: the reference's push task (envs/shape_elasto_plastic.py: ShapeRopeEnv.step_diff = focus
    shift, get_primitive_actions, 20 sub-actions x 16 substeps, reward e^(-10 l2) + e^(-contact)) on a reduced
    particle density, forward + gradient of the summed reward w.r.t. the 6-vector action."""
    sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    import stubs
    stubs.install()
    import torch
    import jax
    import jax.numpy as jnp
    if F64:
        jax.set_float(torch.float64)
    from daxbench.core.envs import shape_elasto_plastic as sep
    from daxbench.core.engine.primitives.primitives import set_sdf
    from daxbench.core.engine.primitives.box import _sdf_batch as box_sdf
    rng = np.random.RandomState(seed)
    env = sep.ShapeRopeEnv(batch_size=B, seed=1, aux_reward=True)
    env.aux_reward = True
    set_sdf(box_sdf)
    conf = env.conf
    # reset() of :139-157 with a reduced density (the shipped density 3 gives 23 940 particles)
    state = env.simulator.add_box(conf=conf, state=None, hardness=conf.rope_hardness, size=conf.rope_width,
                                  init_pos=conf.rope_init_pos, z_rotation_angle=conf.rope_z_rotation_angle, material=2,
                                  density=density)
    state = env.create_primitive(conf=conf, state=state, friction=0.1, color=[0.5, 0.5, 0.5], size=[0.015, 0.06, 0.015],
                                 init_pos=[0.5, 0.01, 0.45])
    env.initialize_after_adding_particle_primitives(state)
    st = env.state
    n = st.x.shape[1]
    goal = np.asarray(st.x)[0] + np.array([0.03, 0.0, 0.02], np.float32)
    env.goal = jnp.array(goal)
    x0 = np.asarray(st.x)
    acts = np.stack([np.concatenate([x0[b].mean(0) + [-0.06, 0, -0.03 * (b + 1)], x0[b].mean(0) + [0.05, 0, 0.02]])
                     for b in range(B)]).astype(np.float32)
    out = {"goal": goal, "actions": acts, "density": np.array(density), "in_x": x0,
           "material": np.asarray(env.simulator.material).astype(np.int32), "h": np.asarray(env.simulator.h).astype(np.float32)}
    if F64:
        st, acts = _f64_override(name, st, jnp)
        env.goal = jnp.array(torch.from_numpy(np.load(os.path.join(GOLD, f"ref_mpmenv_{name}.npz"))["goal"]).double())
        out = {}
    a = torch.from_numpy(acts).requires_grad_(True)
    obs, reward, done, info = env.step_diff(jax.Array(a), st)
    ns = info["state"]
    (ga,) = torch.autograd.grad(reward.t.sum(), [a])
    out.update({"reward": reward.t.detach().numpy(), "g_actions": ga.numpy(), "out_x": np.asarray(ns.x.t.detach()),
                "out_v": np.asarray(ns.v.t.detach()), "out_F": np.asarray(ns.F.t.detach()),
                "out_prim_pos": np.asarray(ns.primitives[0].position.t.detach()), "obs": np.asarray(obs.t.detach())})
    path = os.path.join(GOLD, f"ref_mpmenv_{name}{'_f64' if F64 else ''}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: n={n} B={B} reward={out['reward']} |g_actions|max={np.abs(out['g_actions']).max():.3e}")


def task_env_case(name, B, seed, steps=2):  # noqa: C901
    """Env level for the remaining shipped MPM tasks, at their shipped sizes, run from the reference's own env classes:
      whip  envs/whip_rope_env.py  WhipRopeEnv: 67 elastic particles, position-controlled box gripper, get_primitive_actions
            = actions / 50 with zero rotation, ONE sub-action of 70 substeps per env step;
      rope  envs/shape_rope_env.py ShapeRopeEnv: 582 plastic particles, box pusher, 30 sub-actions x 133 substeps per env
            step, reset() followed by the reference's two random pushes (the fixture carries the pushed state);
      pour  envs/pour_water_env.py PourWaterEnv: 702 liquid particles, two bowl colliders (container SDF),
            get_primitive_actions = [actions / 500, 0 for the second bowl], one sub-action of 23 substeps.
    reset() (with the random xz shift of auto_reset), `steps` consecutive step_diff calls (focus shift, reward
    e^(-10 l2) against goals/<task>/goal.npy), gradient of the summed rewards w.r.t. all actions."""
    sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    import stubs
    stubs.install()
    import torch
    import jax
    import jax.numpy as jnp
    if F64:
        jax.set_float(torch.float64)
    rng = np.random.RandomState(seed)
    if name == "rope":
        from daxbench.core.envs import shape_rope_env as mod
        env = mod.ShapeRopeEnv(batch_size=B, seed=1)
        np.random.seed(seed)                                        # reset() ends with two np.random pushes (:173)
        acts = None
    elif name == "whip":
        from daxbench.core.envs import whip_rope_env as mod
        env = mod.WhipRopeEnv(batch_size=B, seed=1)
        acts = np.concatenate([rng.uniform(-0.8, 0.8, (steps, B, 3)), np.zeros((steps, B, 3))], axis=-1).astype(np.float32)
        acts[:, :, 1] = np.abs(acts[:, :, 1])                       # lift, do not push into the ground
    else:
        from daxbench.core.envs import pour_water_env as mod
        # rendering only (writes a .ply through the `sdf` + trimesh + pyrender stack, stubbed here): not on the path
        mod.PourWaterEnv.create_mesh_for_render = lambda self, size: None
        env = mod.PourWaterEnv(batch_size=B, seed=1)
        acts = rng.uniform(-1.0, 1.0, (steps, B, 6)).astype(np.float32)
        acts[:, :, 3:] *= 20.0                                      # tilt the bowl
    obs, st = env.reset(env.simulator.key_global)
    if acts is None:                                                # push across the rope: start / end around its centroid
        c = np.asarray(st.x).mean(1)
        acts = np.stack([np.concatenate([c + [-0.05, 0, -0.04 * (1 + t)], c + [0.06, 0, 0.05]], axis=1)
                         for t in range(steps)]).astype(np.float32)
    out = {"goal": np.asarray(env.goal), "actions": acts, "in_x": np.asarray(st.x),
           "in_v": np.asarray(st.v), "in_C": np.asarray(st.C), "in_F": np.asarray(st.F), "in_J": np.asarray(st.J),
           "material": np.asarray(env.simulator.material).astype(np.int32), "h": np.asarray(env.simulator.h).astype(np.float32)}
    for q, prim in enumerate(st.primitives):
        out[f"in_prim{q}_pos"] = np.asarray(prim.position)
        out[f"in_prim{q}_rot"] = np.asarray(prim.rotation)
    if F64:
        st, acts = _f64_override(name, st, jnp)
        out = {}
    a = torch.from_numpy(acts).requires_grad_(True)
    total, s = 0, st
    for t in range(steps):
        obs, reward, done, info = env.step_diff(jax.Array(a[t]), s)
        s = info["state"]
        total = total + reward.t.sum()
        out[f"reward{t}"] = reward.t.detach().numpy()
        out[f"x{t}"] = np.asarray(s.x.t.detach())
        out[f"v{t}"] = np.asarray(s.v.t.detach())
        for q, prim in enumerate(s.primitives):
            out[f"prim{q}_pos{t}"] = np.asarray(prim.position.t.detach())
            out[f"prim{q}_rot{t}"] = np.asarray(prim.rotation.t.detach())
    (ga,) = torch.autograd.grad(total, [a])
    out["g_actions"] = ga.numpy()
    out["obs"] = np.asarray(obs.t.detach())
    path = os.path.join(GOLD, f"ref_mpmenv_{name}{'_f64' if F64 else ''}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: n={st.x.shape[1]} B={B} rewards={[out[f'reward{t}'] for t in range(steps)]} "
          f"|g_actions|max={np.abs(out['g_actions']).max():.3e}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--real-jax", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--f64", action="store_true", help="mpmenv_* cases only: the fp64 floor files (see F64 above)")
    args = ap.parse_args()
    global F64
    F64 = args.f64
    mods = load_reference(args.real_jax)
    os.makedirs(GOLD, exist_ok=True)
    cases = {
        # name: (material, n_prim, pos_control, sdf_kind, steps, seed, cot_scale, kwargs)
        "plastic_box": lambda: mpm_case(mods, "plastic_box", 2, 1, False, 0, 4, 11, 1e-3),
        "elastic_2box": lambda: mpm_case(mods, "elastic_2box", 1, 2, False, 0, 4, 12, 1e-3),
        "liquid_bowl": lambda: mpm_case(mods, "liquid_bowl", 0, 2, False, 1, 4, 13, 1e-3, ylow=0.045, v_scale=0.05,
                                        c_scale=1.0, f_scale=0.02),
        "elastic_poscontrol": lambda: mpm_case(mods, "elastic_poscontrol", 1, 1, True, 0, 4, 14, 1e-3),
        "plastic_normgrad": lambda: mpm_case(mods, "plastic_normgrad", 2, 1, False, 0, 4, 15, 10.0),
        "plastic_zero_rot": lambda: mpm_case(mods, "plastic_zero_rot", 2, 1, False, 0, 3, 17, 1e-3, zero_rot=True),
        "plastic_fwd16": lambda: mpm_case(mods, "plastic_fwd16", 2, 1, False, 0, 16, 16, 1e-3, grads=False),
        "cloth_lifted_int": lambda: cloth_case(mods, "lifted_int", False, 21, False),
        "cloth_lifted_float": lambda: cloth_case(mods, "lifted_float", True, 22, False),
        "cloth_reset_grab": lambda: cloth_case(mods, "reset_grab", True, 23, True),
        "cloth_w1_reset": lambda: cloth_case(mods, "w1_reset", True, 24, True, window=1),
        "cloth_w3_reset": lambda: cloth_case(mods, "w3_reset", False, 25, True, window=3),
        "cloth_w5_lifted": lambda: cloth_case(mods, "w5_lifted", True, 26, False, window=5),
        "cloth_w1_contact": lambda: cloth_case(mods, "w1_contact", True, 27, False, window=1, contact=True),
        "cloth_w4_contact": lambda: cloth_case(mods, "w4_contact", False, 28, False, window=4, contact=True),
    }
    cases["mpmenv_push"] = lambda: mpm_env_case("push", 2, 1.3, 41)
    cases["mpmenv_whip"] = lambda: task_env_case("whip", 2, 43)
    cases["mpmenv_pour"] = lambda: task_env_case("pour", 2, 44)
    cases["mpmenv_rope"] = lambda: task_env_case("rope", 1, 45, steps=1)      # 3 990 substeps per env step: ~30 min here
    cases["clothenv_unfold1"] = lambda: cloth_unfold_case("unfold1", 2, 51)
    cases["clothenv_ep1"] = lambda: cloth_env_case("ep1", 1, 2, 31)
    # BASELINE.json configs[3]: fold_cloth1_para, stiffness of training iteration 3 (apg_para.py:326-329)
    cases["clothenv_para"] = lambda: cloth_env_para_case("para", 2, 33, 3)
    # BASELINE.json configs[0]: fold_cloth3 APG ep_len=3 num_envs=4 (reference states + policy gradient; ~20 min here)
    cases["clothenv_ep3"] = lambda: cloth_env_case("ep3", 3, 4, 0)
    for name, fn in cases.items():
        if args.only and args.only not in name:
            continue
        if F64 and not name.startswith("mpmenv_"):
            continue
        fn()


if __name__ == "__main__":
    main()
