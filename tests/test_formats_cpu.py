"""unidom_b200.formats: the reference's pickles (expert demos, APG policy checkpoints) read without jax / flax.
The pickles the reference ships are read where /root/reference exists (this container); a self-contained case builds a
checkpoint with stand-in classes that pickle the way jax 0.3.14's DeviceArray and flax's FrozenDict do."""
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from unidom_b200 import apg, formats

REF = "/root/reference/DaXBench/daxbench"


def _fake_jax_modules():
    """Stand-ins that reduce like the real classes: DeviceArray -> (reconstruct_device_array, (numpy reduce..., aval)),
    FrozenDict -> object state {"_dict": ...}."""
    jmod = types.ModuleType("jax._src.device_array")

    def reconstruct_device_array(fun, args, arr_state, aval_state):
        raise AssertionError("must not be called: the loader rebuilds the NumPy array itself")
    reconstruct_device_array.__module__, reconstruct_device_array.__qualname__ = "jax._src.device_array", "reconstruct_device_array"
    jmod.reconstruct_device_array = reconstruct_device_array

    class DeviceArray:
        def __init__(self, v):
            self.v = np.asarray(v)

        def __reduce__(self):
            fun, args, arr_state = self.v.__reduce__()
            return (reconstruct_device_array, (fun, args, arr_state, {"weak_type": False, "named_shape": {}}))
    fmod = types.ModuleType("flax.core.frozen_dict")

    class FrozenDict:
        def __init__(self, d):
            self._dict = d
            self._hash = None
    FrozenDict.__module__, FrozenDict.__qualname__ = "flax.core.frozen_dict", "FrozenDict"
    fmod.FrozenDict = FrozenDict
    return {"jax": types.ModuleType("jax"), "jax._src": types.ModuleType("jax._src"), "jax._src.device_array": jmod,
            "flax": types.ModuleType("flax"), "flax.core": types.ModuleType("flax.core"), "flax.core.frozen_dict": fmod}, DeviceArray, FrozenDict


def test_policy_checkpoint_round_trips_through_the_reference_layout(tmp_path):
    params = apg.init_policy(1544, 8, seed=3)
    mods, DA, FD = _fake_jax_modules()
    tree = FD({"params": FD({f"hidden_{i // 2}": FD({"kernel": DA(params[i].numpy()), "bias": DA(params[i + 1].numpy())})
                             for i in range(0, 6, 2)})})
    saved = dict(sys.modules)
    sys.modules.update(mods)
    try:
        path = tmp_path / "apg_fold_cloth3_0.pkl"
        with open(path, "wb") as f:
            pickle.dump(tree, f)
    finally:
        for k in mods:
            sys.modules.pop(k, None)
        sys.modules.update({k: v for k, v in saved.items() if k in mods})
    got = formats.load_policy(str(path))
    assert len(got) == 6
    for a, b in zip(got, params):
        assert a.dtype == torch.float32 and torch.equal(a, b)
    # and what this package writes is readable by plain pickle and by its own loader
    out = tmp_path / "mine.pkl"
    formats.save_policy(params, str(out))
    plain = pickle.load(open(out, "rb"))
    assert sorted(plain["params"]) == ["hidden_0", "hidden_1", "hidden_2"]
    assert plain["params"]["hidden_2"]["kernel"].shape == (256, 16)
    for a, b in zip(formats.load_policy(str(out)), params):
        assert torch.equal(a, b)
    obs = torch.randn(4, 1544)
    assert torch.equal(apg.policy_apply(got, obs), apg.policy_apply(params, obs))


@pytest.mark.skipif(not os.path.exists(f"{REF}/algorithms/expert_demo/fold_cloth3/demo_0.pkl"), reason="reference tree absent")
def test_reference_expert_demos_load_into_this_packages_states():
    from unidom_b200.cloth_simulator import ClothState
    from unidom_b200.mpm_simulator import MPMState, PrimitiveState
    d = formats.load_demo(f"{REF}/algorithms/expert_demo/fold_cloth3/demo_0.pkl")
    assert len(d["action"]) == len(d["state"]) or len(d["action"]) + 1 == len(d["state"])
    s0 = d["state"][0]
    assert isinstance(s0, ClothState) and s0.x.dtype == torch.float32 and s0.x.shape[-1] == 3
    assert d["action"][0].numel() in (6, 8)       # env-level pick-and-place action
    m = formats.load_demo(f"{REF}/algorithms/expert_demo/whip_rope/demo_0.pkl")
    ms = m["state"][0]
    assert isinstance(ms, MPMState) and isinstance(ms.primitives[0], PrimitiveState)
    assert ms.F.shape[-2:] == (3, 3) and ms.primitives[0].position.shape[-1] == 3
