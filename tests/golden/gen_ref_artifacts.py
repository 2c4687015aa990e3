#!/usr/bin/env python
"""Copy the small reference ARTEFACTS used as known-answer tests into tests/golden/ref_artifacts.npz
(/root/reference does not travel to the GPU box):
  * core/envs/goals/shape_rope/goal.npy -- equals the deterministic add_box lattice of shape_rope's reset
    (envs/shape_rope_env.py:162-164 -> mpm_simulator.py:93-109);
  * algorithms/expert_demo/fold_cloth3/demo_0.pkl -- actions + gripper (primitive0) states per env step: pins
    get_pnp_actions (cloth_env.py:136-173) and the /50 action scaling + clip of robot_step (cloth_simulator.py:168);
  * algorithms/expert_demo/whip_rope/demo_0.pkl -- primitive position per env step vs the commanded displacement:
    pins the FK row-S write-drop / read-clamp rule (primitives.py:185-194) and reset's Lame parameters.
The pickles hold jax DeviceArrays; they are read with a stub unpickler (no jax needed)."""
import os
import pickle

import numpy as np

REF = "/root/reference/DaXBench/daxbench"
HERE = os.path.dirname(os.path.abspath(__file__))


class _T(tuple):
    def __new__(cls, *a):
        return tuple.__new__(cls, a)


class Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name in ("ClothState", "MPMState", "PrimitiveState"):
            return _T
        if module.startswith("jax"):
            def rebuild(fun, args, state, *rest):
                a = fun(*args)
                a.__setstate__(state)
                return a
            return rebuild
        if module == "numpy.core.multiarray":
            import numpy._core.multiarray as m
            return getattr(m, name)
        if module == "numpy.core.numeric":
            import numpy._core.numeric as m
            return getattr(m, name)
        return super().find_class(module, name)


def load(path):
    with open(path, "rb") as f:
        return Unpickler(f).load()


def main():
    out = {"shape_rope_goal": np.load(f"{REF}/core/envs/goals/shape_rope/goal.npy")}
    d = load(f"{REF}/algorithms/expert_demo/fold_cloth3/demo_0.pkl")
    # ClothState field order: x v primitive0 primitive1 action0 action1 key cur_step stiffness mu
    out["cloth_actions"] = np.stack([np.asarray(a) for a in d["action"]]).astype(np.float32)
    out["cloth_primitive0"] = np.stack([np.asarray(s[2]) for s in d["state"]]).astype(np.float32)
    out["cloth_x0"] = np.asarray(d["state"][0][0]).astype(np.float32)
    d = load(f"{REF}/algorithms/expert_demo/whip_rope/demo_0.pkl")
    # MPMState: x v C F J cur_step primitives key friction mu lamda ; PrimitiveState: size dim friction softness
    # color position rotation v w xyz_limit action_buffer action_scale ...
    out["whip_actions"] = np.stack([np.asarray(a) for a in d["action"]]).astype(np.float32)
    out["whip_prim_pos0"] = np.stack([np.asarray(s[6][0][5])[0, 0] for s in d["state"]]).astype(np.float32)
    out["whip_prim_action_scale"] = np.asarray(d["state"][0][6][0][11]).astype(np.float32)
    out["whip_prim_steps"] = np.array(np.asarray(d["state"][0][6][0][5]).shape[1])
    out["whip_mu_lamda"] = np.array([np.asarray(d["state"][0][9]).ravel()[0], np.asarray(d["state"][0][10]).ravel()[0]], np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_artifacts.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
