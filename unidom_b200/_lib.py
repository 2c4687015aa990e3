"""ctypes binding of libunidom_b200.so (the C ABI in include/unidom_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libunidom_b200.so")

UD_MAX_PRIM = 4
UD_SDF_BOX, UD_SDF_CONTAINER = 0, 1
UD_P2G_ATOMIC, UD_P2G_DETERMINISTIC = 0, 1
UD_P2G_LIQUID_FAST = 0x100

_fp = C.c_void_p  # device pointers travel as integers


class MpmParams(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("n_particles", C.c_int32), ("steps", C.c_int32),
        ("res", C.c_int32 * 3), ("n_grid", C.c_int32),
        ("dt", C.c_double), ("dx", C.c_double), ("inv_dx", C.c_double),
        ("p_mass", C.c_double), ("p_vol", C.c_double), ("gravity", C.c_double * 3),
        ("n_primitive", C.c_int32), ("sdf_kind", C.c_int32),
        ("use_position_control", C.c_int32), ("p2g_mode", C.c_int32),
    ]


PRIM_FIELDS = ("size", "friction", "softness", "position", "rotation", "v", "w", "action_buffer", "action_scale")


class Primitive(C.Structure):
    _fields_ = [(k, _fp) for k in PRIM_FIELDS]


STATE_FIELDS = ("x", "v", "C", "F", "J", "friction", "mu", "lamda")


class MpmState(C.Structure):
    _fields_ = [(k, _fp) for k in STATE_FIELDS] + [("prim", Primitive * UD_MAX_PRIM)]


class ClothParams(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("n_nodes", C.c_int32), ("N", C.c_int32), ("substeps", C.c_int32),
        ("dt", C.c_double), ("gravity", C.c_double), ("damping", C.c_double), ("max_v", C.c_double),
        ("small_num", C.c_double), ("cell_size", C.c_double), ("mask_sum", C.c_double),
        ("stiffness_is_float", C.c_int32),
    ]


CLOTH_FIELDS = ("x", "v", "primitive0", "primitive1", "action0", "action1", "stiffness", "mu")


class ClothState(C.Structure):
    _fields_ = [(k, _fp) for k in CLOTH_FIELDS]


_lib = None


def lib():
    """Load (once) and return the shared library; raise loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UNIDOM_B200_LIB") or LIB_PATH      # development A/B builds (build.py --variant)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m unidom_b200.build` "
            "(nvcc, sm_100a).  unidom_b200 has no CPU or eager fallback.")
    L = C.CDLL(path)
    L.ud_version.restype = C.c_char_p
    L.ud_last_error.restype = C.c_char_p
    P = C.POINTER
    L.ud_mpm_fwd_workspace_bytes.restype = C.c_size_t
    L.ud_mpm_fwd_workspace_bytes.argtypes = [P(MpmParams)]
    L.ud_mpm_bwd_workspace_bytes.restype = C.c_size_t
    L.ud_mpm_bwd_workspace_bytes.argtypes = [P(MpmParams)]
    L.ud_mpm_num_keys.restype = C.c_int32
    L.ud_mpm_num_keys.argtypes = [P(MpmParams)]
    L.ud_mpm_step_fwd.restype = C.c_int
    L.ud_mpm_step_fwd.argtypes = [P(MpmParams), P(MpmState), _fp, _fp, _fp, P(MpmState), _fp, C.c_size_t, _fp]
    L.ud_mpm_step_bwd.restype = C.c_int
    L.ud_mpm_step_bwd.argtypes = [P(MpmParams), P(MpmState), _fp, _fp, _fp, P(MpmState), P(MpmState), _fp,
                                  _fp, C.c_size_t, _fp]
    L.ud_mpm_bwd_windowed_workspace_bytes.restype = C.c_size_t
    L.ud_mpm_bwd_windowed_workspace_bytes.argtypes = [P(MpmParams), C.c_int32]
    L.ud_mpm_step_bwd_windowed.restype = C.c_int
    L.ud_mpm_step_bwd_windowed.argtypes = [P(MpmParams), P(MpmState), _fp, _fp, _fp, P(MpmState), P(MpmState), _fp,
                                           C.c_int32, _fp, C.c_size_t, _fp]
    L.ud_mpm_tape_bytes.restype = C.c_size_t
    L.ud_mpm_tape_bytes.argtypes = [P(MpmParams)]
    L.ud_mpm_step_fwd_taped.restype = C.c_int
    L.ud_mpm_step_fwd_taped.argtypes = [P(MpmParams), P(MpmState), _fp, _fp, _fp, P(MpmState), _fp, C.c_size_t, _fp]
    L.ud_mpm_step_bwd_taped.restype = C.c_int
    L.ud_mpm_step_bwd_taped.argtypes = [P(MpmParams), P(MpmState), _fp, P(MpmState), P(MpmState), _fp, _fp,
                                        C.c_size_t, _fp]
    i32 = C.c_int32
    L.ud_chamfer_residual_bytes.restype = C.c_size_t
    L.ud_chamfer_residual_bytes.argtypes = [i32, i32, i32]
    L.ud_chamfer_fwd.restype = C.c_int
    L.ud_chamfer_fwd.argtypes = [_fp, _fp, i32, i32, i32, _fp, _fp, C.c_size_t, _fp]
    L.ud_chamfer_bwd.restype = C.c_int
    L.ud_chamfer_bwd.argtypes = [_fp, _fp, i32, i32, i32, _fp, _fp, C.c_size_t, _fp, _fp]
    L.ud_l2_fwd.restype = C.c_int
    L.ud_l2_fwd.argtypes = [_fp, _fp, i32, i32, _fp, _fp]
    L.ud_l2_bwd.restype = C.c_int
    L.ud_l2_bwd.argtypes = [_fp, _fp, i32, i32, _fp, _fp, _fp]
    L.ud_apg_scrub_clip.restype = C.c_int
    L.ud_apg_scrub_clip.argtypes = [_fp, C.c_int64, C.c_float, _fp, _fp]
    L.ud_adam_step.restype = C.c_int
    L.ud_adam_step.argtypes = [_fp, _fp, _fp, _fp, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double,
                               C.c_int32, _fp]
    L.ud_apg_fused_update.restype = C.c_int
    L.ud_apg_fused_update.argtypes = [_fp, _fp, _fp, _fp, C.c_int64, C.c_float, C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.c_int32, C.c_int32, C.c_int32, _fp, _fp, _fp, _fp]
    L.ud_mpm_sort_bins.restype = C.c_int
    L.ud_mpm_sort_bins.argtypes = [P(MpmParams), _fp, _fp, _fp, _fp, _fp, C.c_size_t, _fp]
    if hasattr(L, "ud_cloth_step_fwd"):
        L.ud_cloth_workspace_bytes.restype = C.c_size_t
        L.ud_cloth_workspace_bytes.argtypes = [P(ClothParams)]
        L.ud_cloth_step_fwd.restype = C.c_int
        L.ud_cloth_step_fwd.argtypes = [P(ClothParams), P(ClothState), _fp, _fp, _fp, P(ClothState), _fp, C.c_size_t, _fp]
        L.ud_cloth_step_bwd.restype = C.c_int
        L.ud_cloth_step_bwd.argtypes = [P(ClothParams), P(ClothState), _fp, _fp, _fp, P(ClothState), P(ClothState),
                                        _fp, _fp, C.c_size_t, _fp]
    L.ud_cloth_multi_ckpt_bytes.restype = C.c_size_t
    L.ud_cloth_multi_ckpt_bytes.argtypes = [P(ClothParams), C.c_int32]
    L.ud_cloth_multi_workspace_bytes.restype = C.c_size_t
    L.ud_cloth_multi_workspace_bytes.argtypes = [P(ClothParams), C.c_int32]
    L.ud_cloth_multi_step_fwd.restype = C.c_int
    L.ud_cloth_multi_step_fwd.argtypes = [P(ClothParams), P(ClothState), _fp, _fp, _fp, C.c_int32, P(ClothState), _fp,
                                          C.c_size_t, _fp]
    L.ud_cloth_multi_step_bwd.restype = C.c_int
    L.ud_cloth_multi_step_bwd.argtypes = [P(ClothParams), P(ClothState), _fp, _fp, _fp, C.c_int32, _fp, P(ClothState),
                                          P(ClothState), _fp, _fp, C.c_size_t, _fp]
    for kv in filter(None, os.environ.get("UNIDOM_B200_TUNE", "").split(",")):   # development A/B switches, "warp=2,..."
        name, val = kv.split("=")
        if L.ud_tuning_set(name.encode(), int(val)) < 0:
            raise ValueError(f"UNIDOM_B200_TUNE: unknown switch {kv!r}")
    _lib = L
    return L


EXPORTS = (
    "ud_version", "ud_last_error", "ud_mpm_fwd_workspace_bytes", "ud_mpm_bwd_workspace_bytes",
    "ud_mpm_step_fwd", "ud_mpm_step_bwd", "ud_mpm_bwd_windowed_workspace_bytes", "ud_mpm_step_bwd_windowed",
    "ud_mpm_tape_bytes", "ud_mpm_step_fwd_taped", "ud_mpm_step_bwd_taped",
    "ud_mpm_sort_bins", "ud_mpm_num_keys",
    "ud_cloth_workspace_bytes", "ud_cloth_step_fwd", "ud_cloth_step_bwd", "ud_cloth_multi_ckpt_bytes",
    "ud_cloth_multi_workspace_bytes", "ud_cloth_multi_step_fwd", "ud_cloth_multi_step_bwd",
    "ud_chamfer_residual_bytes", "ud_chamfer_fwd", "ud_chamfer_bwd", "ud_l2_fwd", "ud_l2_bwd", "ud_apg_scrub_clip", "ud_adam_step", "ud_apg_fused_update",
    "ud_launch_count", "ud_timing_enable", "ud_timing_collect", "ud_tuning_set",
)


def check(code, what):
    if code != 0:
        msg = lib().ud_last_error().decode()
        raise RuntimeError(f"{what} failed with status {code}: {msg}")
