"""Oracle parity AT BASELINE.json's particle counts: one env of configs[1] (push_plasticine, 50 625 plastic particles,
59 particles per cell, S = 16), of configs[2] (pour_water, 99 998 liquid particles, two bowl colliders, S = 23) and of
configs[4] (whip_rope at density 25, 49 329 elastic particles, position control, S = 70) -- forward state and adjoint of
one `step_jax` call from a settled mid-episode state, CUDA (through the C ABI) against the CPU oracle in fp32, with
the oracle's own fp32-vs-fp64 floor printed next to every error.

The small-scene tests never reach the regime the bench runs in (dozens of particles per cell, 4-6 cells per warp,
tens of thousands of CTAs); these do.  Bars (BASELINE.json north_star): state rtol 1e-4, gradients rtol 1e-3 with
cosine >= 0.999 -- or the printed fp32 floor of the reference arithmetic itself where that is higher.
"""
import numpy as np
import pytest
import torch

import util
from oracle import mpm as omp

pytestmark = pytest.mark.gpu

STATE_BAR, GRAD_BAR, COS_BAR = 1e-4, 1e-3, 0.999


def _oracle(conf, sim, dtype):
    return omp.Simulator(util.oracle_conf(conf), sim.material.clone(), sim.h.clone(), dtype=dtype, checkpoint_substeps=True)


def _grads(step_fn, state, action, cot, to_dev):
    names = ("x", "v", "C", "F", "mu", "lamda", "friction")
    req = {k: getattr(state, k).detach().clone().requires_grad_(True) for k in names}
    a = action.detach().clone().requires_grad_(True)
    out = step_fn(state._replace(**req), a)
    L = sum((getattr(out, k) * to_dev(cot[k]).to(getattr(out, k).dtype)).sum() for k in cot)
    gs = torch.autograd.grad(L, [req[k] for k in names] + [a], allow_unused=True)
    res = {k: (g if g is not None else torch.zeros_like(req[k])) for k, g in zip(names, gs[:-1])}
    res["action"] = gs[-1]
    return out, res


def _check(label, conf, sim, st, act, cot_scale=1e-3, seed=0):
    dev = st.x.device
    B, n = st.x.shape[:2]
    g = torch.Generator().manual_seed(seed)
    cot = {"x": torch.randn((B, n, 3), generator=g) * cot_scale, "v": torch.randn((B, n, 3), generator=g) * cot_scale * 0.1,
           "C": torch.randn((B, n, 3, 3), generator=g) * cot_scale * 1e-3, "F": torch.randn((B, n, 3, 3), generator=g) * cot_scale * 0.1}
    out, got = _grads(lambda s, a: sim.step_jax(s, a)[0], st, act, cot, lambda t: t.to(dev))
    o32, r32 = _grads(lambda s, a: omp.step_batch(_oracle(conf, sim, torch.float32), s, a), util.to_oracle_state(st),
                      act.cpu(), cot, lambda t: t)
    o64, r64 = _grads(lambda s, a: omp.step_batch(_oracle(conf, sim, torch.float64), s, a),
                      util.to_oracle_state(st, torch.float64), act.cpu().double(), {k: v.double() for k, v in cot.items()},
                      lambda t: t)
    ok = True
    for k in ("x", "v", "C", "F", "J"):
        e, fl = util.rel_err(getattr(out, k), getattr(o32, k)), util.rel_err(getattr(o32, k), getattr(o64, k))
        e64 = util.rel_err(getattr(out, k), getattr(o64, k))
        bar = max(STATE_BAR, 3 * fl)
        print(f"[{label}] state {k}: cuda-vs-oracle32 {e:.3e}  cuda-vs-oracle64 {e64:.3e}  oracle32-vs-oracle64 (floor) {fl:.3e}  bar {bar:.1e}")
        ok &= e < bar
    for k in r32:
        if float(r32[k].abs().max()) <= 1e-12:
            continue
        e, fl = util.rel_err(got[k], r32[k]), util.rel_err(r32[k], r64[k])
        cs, cs64 = util.cosine(got[k], r32[k]), util.cosine(got[k], r64[k])
        bar = max(GRAD_BAR, 20 * fl)
        print(f"[{label}] grad {k:8s}: rel {e:.3e} cos {cs:.7f} (vs fp64 oracle cos {cs64:.7f})  floor {fl:.3e}  bar {bar:.1e}  "
              f"max|ref| {float(r32[k].abs().max()):.3e}")
        ok &= cs >= COS_BAR and e < bar
    assert ok, label


def test_push_plasticine_50625_particles(built_lib):
    """configs[1]: the bench scene itself (tests/test_fullsize_gpu.py::_scene), ONE env, after 4 pushing steps."""
    from unidom_b200 import confs
    from unidom_b200.mpm_simulator import SimpleMPMSimulator, create_primitive
    conf = confs.shape_elasto_plastic_conf()
    sim = SimpleMPMSimulator(conf, 1)
    st = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.2, 0.06, 0.12], init_pos=[0.25, 0.07, 0.25],
                     z_rotation_angle=0, material=2, density=3.9)
    st.primitives.append(create_primitive(conf, friction=0.1, softness=666, color=[0.5] * 3, size=[0.015, 0.06, 0.015],
                                          init_pos=[0.25, 0.01, 0.20]))
    st = st._replace(primitives=[p._replace(action_scale=p.action_scale * 0.012) for p in st.primitives])
    st = sim.reset_jax(st)
    assert st.x.shape[1] == 50625
    act = torch.tensor([[0.25, 0.0, 0.33, 0.0, 0.0, 0.0]], device="cuda")
    with torch.no_grad():
        for _ in range(4):
            st, _ = sim.step_jax(st, act)
    ppc = st.x.shape[1] / len(torch.unique((st.x[0] * conf.inv_dx - 0.5).int(), dim=0))
    print(f"[push] particles per occupied cell {ppc:.1f}")
    _check("push_plasticine 50625", conf, sim, st, act)


def test_pour_water_99998_particles(built_lib):
    """configs[2]: SURVEY 8d's synthetic pour_water env (cube side 0.3655 at density 4, res (64,48,64), two bowls)."""
    from unidom_b200 import _lib, confs
    from unidom_b200.mpm_simulator import SimpleMPMSimulator, create_primitive
    conf = confs.pour_water_conf(res=(64, 48, 64))
    sim = SimpleMPMSimulator(conf, 1, sdf_kind=_lib.UD_SDF_CONTAINER)
    st = sim.add_box(conf=conf, state=None, hardness=1, size=[0.3655] * 3, init_pos=[0.4, 0.3, 0.4], material=0, density=4)
    st.primitives.append(create_primitive(conf, 0.1, 666, [0.5] * 3, [0.35, 0.0, 0.02], [0.4, 0.3, 0.4]))
    st.primitives.append(create_primitive(conf, 0.1, 666, [0.5] * 3, [0.3, 0.0, 0.02], [0.4, 0.08, 0.2]))
    st = st._replace(primitives=[p._replace(action_scale=p.action_scale * 0.01) for p in st.primitives])
    st = sim.reset_jax(st)
    assert st.x.shape[1] == 99998
    act = torch.zeros((1, 12), device="cuda")
    act[:, 0], act[:, 5] = 0.3, 0.2
    with torch.no_grad():
        for _ in range(3):
            st, _ = sim.step_jax(st, act)
    _check("pour_water 99998", conf, sim, st, act)


def test_whip_rope_49329_particles(built_lib):
    """configs[4]: the whip_rope scene (envs/whip_rope_env.py) at density 25, S = 70 substeps, position control."""
    from unidom_b200 import confs
    from unidom_b200.mpm_simulator import SimpleMPMSimulator
    conf = confs.whip_rope_conf()
    assert conf.steps == 70
    sim = SimpleMPMSimulator(conf, 1, use_position_control=True)
    st = confs.build_whip_rope(sim, density=25)
    assert st.x.shape[1] == 49329
    st = st._replace(primitives=[p._replace(action_scale=p.action_scale * 0.02) for p in st.primitives])
    act = torch.tensor([[0.2, 0.5, 0.1, 0.0, 0.0, 0.0]], device="cuda")
    with torch.no_grad():
        st, _ = sim.step_jax(st, act)
    _check("whip_rope 49329", conf, sim, st, act)
