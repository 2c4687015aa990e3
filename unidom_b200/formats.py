"""The reference's on-disk formats, readable without jax / flax:

  expert demos      core/envs/basic/cloth_env.py:286-318 (and mpm_env): pickle of {"action": [...], "state": [...], ...}
                    whose leaves are jax DeviceArrays inside ClothState / MPMState / PrimitiveState NamedTuples
  policy checkpoint algorithms/apg/apg.py:325-330: pickle of the flax parameter tree of the 512-256-2A MLP
                    {"params": {"hidden_0": {"kernel", "bias"}, "hidden_1": ..., "hidden_2": ...}} (device 0's replica)

A DeviceArray pickles as (reconstruct, (numpy's own reduce tuple, ...)); the unpickler below rebuilds the NumPy array and
maps the reference's NamedTuples onto this package's, so a demo recorded by the reference replays through
`ClothEnv.step_diff` and a policy trained there initialises `apg.policy_apply` (and the other way round: checkpoints
are written as plain dicts of NumPy arrays, which flax accepts as a parameter tree)."""
import pickle

import numpy as np
import torch


def _state_classes():
    from .cloth_simulator import ClothState
    from .mpm_simulator import MPMState, PrimitiveState
    return {"ClothState": ClothState, "MPMState": MPMState, "PrimitiveState": PrimitiveState}


class _FrozenDict(dict):
    """flax.core.frozen_dict.FrozenDict pickles as an object whose state is {"_dict": {...}} (+ a cached hash)."""

    def __setstate__(self, state):
        self.update(state.get("_dict", state))


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        cls = _state_classes().get(name)
        if cls is not None:
            return cls                            # same field order as the reference's NamedTuples (NEWOBJ needs a type)
        if name == "FrozenDict":
            return _FrozenDict
        if module.startswith("jax"):            # jax._src.device_array.reconstruct_device_array(fun, args, arr_state, aval_state)
            def rebuild(fun, args, state, *rest):
                a = fun(*args)
                a.__setstate__(state)
                return a
            return rebuild
        if module.startswith("numpy.core"):     # pickles written by numpy < 2
            import importlib
            return getattr(importlib.import_module(module.replace("numpy.core", "numpy._core")), name)
        return super().find_class(module, name)


def _to_torch(obj):
    if isinstance(obj, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(obj))
    if isinstance(obj, tuple) and hasattr(obj, "_fields"):
        return type(obj)(*[_to_torch(v) for v in obj])
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_torch(v) for v in obj)
    if isinstance(obj, dict):
        return {k: _to_torch(v) for k, v in obj.items()}
    return obj


def load_pickle(path):
    """Any pickle of the reference (jax arrays -> torch CPU tensors, its NamedTuples -> this package's)."""
    with open(path, "rb") as f:
        return _to_torch(_Unpickler(f).load())


def load_demo(path):
    """expert_demo/<task>/demo_i.pkl -> {"action": [tensor], "state": [ClothState | MPMState], ...}."""
    d = load_pickle(path)
    if not isinstance(d, dict) or "action" not in d or "state" not in d:
        raise ValueError(f"{path}: not an expert demo (expected a dict with 'action' and 'state')")
    return d


def load_policy(path):
    """apg_<env>_<it>.pkl -> [W1, b1, W2, b2, W3, b3] (the layout of unidom_b200.apg)."""
    tree = load_pickle(path)
    tree = tree.get("params", tree)
    layers = sorted((k for k in tree if k.startswith("hidden_")), key=lambda k: int(k.split("_")[1]))
    if not layers:
        raise ValueError(f"{path}: no hidden_<i> layers in the parameter tree ({list(tree)})")
    out = []
    for k in layers:
        out += [tree[k]["kernel"].to(torch.float32), tree[k]["bias"].to(torch.float32)]
    return out


def save_policy(params, path):
    """The inverse: a plain-dict parameter tree of NumPy arrays (loadable by pickle.load without this package, and a
    valid flax parameter tree)."""
    tree = {"params": {f"hidden_{i // 2}": {"kernel": params[i].detach().cpu().numpy(), "bias": params[i + 1].detach().cpu().numpy()}
                       for i in range(0, len(params), 2)}}
    with open(path, "wb") as f:
        pickle.dump(tree, f)
