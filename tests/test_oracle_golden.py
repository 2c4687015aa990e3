"""CPU: the oracle restatement against golden vectors produced by the UNMODIFIED reference sources
(oracle/gen_golden.py runs /root/reference/DaXBench/daxbench/core/engine/*.py under oracle/jaxshim).
This is what pins the oracle: same inputs, the reference's own Python on one side, oracle/ on the other."""
import glob
import os

import numpy as np
import pytest
import torch

import util
from oracle import cloth as oc
from oracle import mpm as omp

MPM_CASES = sorted(os.path.basename(p)[len("ref_mpm_"):-4] for p in glob.glob(os.path.join(util.GOLD, "ref_mpm_*.npz")))
CLOTH_CASES = sorted(os.path.basename(p)[len("ref_cloth_"):-4] for p in glob.glob(os.path.join(util.GOLD, "ref_cloth_*.npz")))


def test_fixtures_present():
    assert len(MPM_CASES) >= 5 and len(CLOTH_CASES) >= 3, (MPM_CASES, CLOTH_CASES)


@pytest.mark.parametrize("name", MPM_CASES)
def test_mpm_oracle_matches_reference_sources(name):
    conf, d = util.golden_mpm(name)
    osim = omp.Simulator(util.oracle_conf(conf), d["material"].to(torch.int32), d["h"])
    st = util.golden_oracle_state(conf, d)
    has_grads = "g_x" in d
    if has_grads:
        got, out = util.golden_mpm_grads(lambda s, a: omp.step_batch(osim, s, a), st, d["action"], d, conf.n_primitive,
                                         lambda t: t)
    else:
        with torch.no_grad():
            out = omp.step_batch(osim, st, d["action"])
    for k in util.STATE_F:
        e = util.rel_err(getattr(out, k), d["out_" + k])
        print(f"{name} state {k}: oracle-vs-reference rel {e:.3e}")
        assert e < (1e-3 if conf.sdf_kind == 1 else (5e-6 if conf.steps > 8 else 2e-6)), (k, e)   # 2 bowls + liquid: 1e-7 for 3 substeps, then one grid cell flips a collider branch (FD normals, d=1e-6 in fp32)
    for q in range(conf.n_primitive):
        for k in util.PRIM_F:
            e = util.rel_err(getattr(out.primitives[q], k), d[f"out_p{q}_{k}"])
            assert e < 1e-6, (q, k, e)
    if not has_grads:
        return
    for k, g in got.items():
        ref = d["g_" + k]
        if ("in_" + k) in d and not d["in_" + k].is_floating_point():
            continue      # integer leaf in the reference (e.g. ground_friction = 2): float0 cotangent, no gradient
        if float(ref.abs().max()) < 1e-20:
            assert float(g.abs().max()) < 1e-12, k
            continue
        e, cs = util.rel_err(g, ref), util.cosine(g, ref)
        print(f"{name} grad {k:16s}: rel {e:.3e} cos {cs:.8f} max|ref| {float(ref.abs().max()):.3e}")
        assert cs > 0.99999 and e < 1e-3, (k, e, cs)


def _cloth_state(d, prefix="in_", dtype=torch.float32):
    B = d[prefix + "x"].shape[0]
    vals = {k: d[prefix + k] for k in ("x", "v", "primitive0", "primitive1", "action0", "action1", "stiffness", "mu")}
    vals = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in vals.items()}
    return oc.ClothState(key=torch.zeros((B, 2), dtype=torch.int32), cur_step=torch.zeros(B, dtype=torch.int32), **vals)


@pytest.mark.parametrize("name", CLOTH_CASES)
def test_cloth_oracle_matches_reference_sources(name):
    """window = K > 0: K substeps of the reference's own step_wrapper (teacher-forced, tight); window = 0: the
    full 50-substep robot_step -- chaotic (SURVEY hard part 2), so only a loose bound is asserted there and
    the numbers are printed."""
    d = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in np.load(os.path.join(util.GOLD, f"ref_cloth_{name}.npz")).items()}
    conf = oc.ClothConf()
    mask = oc.fold_cloth_mask(conf)
    assert np.array_equal(np.asarray(mask).astype(np.int32), d["mask"].numpy())
    osim = oc.ClothSim(conf, mask)
    st = _cloth_state(d)
    window = int(d["window"])
    substeps = window if window else 50
    names = ["x", "v", "primitive0", "primitive1", "mu"] + (["stiffness"] if int(d["stiffness_is_float"]) else [])
    req = {k: getattr(st, k).detach().clone().requires_grad_(True) for k in names}
    a = d["action"].clone().requires_grad_(True)
    s = st._replace(**req)
    for _ in range(int(d["n_calls"])):
        s = oc.step_batch(osim, s, a, substeps)
    for k in ("x", "v", "primitive0", "primitive1", "action0", "action1"):
        e = util.rel_err(getattr(s, k), d["out_" + k])
        print(f"cloth {name} state {k}: oracle-vs-reference rel {e:.3e}")
        assert e < ((2e-6 if window <= 3 else 1e-4) if window else 2e-2), (k, e)   # errors grow ~5x per substep (stiff explicit springs)
    L = sum((getattr(s, k) * d["cot_" + k]).sum() for k in ("x", "v", "primitive0", "primitive1"))
    gr = torch.autograd.grad(L, [req[k] for k in names] + [a], allow_unused=True)
    for k, g in zip(names + ["action"], gr):
        ref = d["g_" + k]
        g = g if g is not None else torch.zeros_like(ref)
        if float(ref.abs().max()) < 1e-20:
            assert float(g.abs().max()) < 1e-12, k
            continue
        e, cs = util.rel_err(g, ref), util.cosine(g, ref)
        print(f"cloth {name} grad {k:12s}: rel {e:.3e} cos {cs:.10f} max|ref| {float(ref.abs().max()):.3e}")
        if window:
            assert cs > (0.999999 if window <= 3 else 0.9999) and e < (1e-4 if window <= 3 else 1e-2), (k, e, cs)
        else:
            assert cs > 0.9, (k, e, cs)
