"""CPU-side checks: the C-ABI library builds for sm_100a, loads, and exports every declared symbol."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols(built_lib):
    hdr = open(os.path.join(ROOT, "include", "unidom_b200.h")).read()
    declared = set(re.findall(r"\b(ud_[a-z0-9_]+)\s*\(", hdr))
    assert {"ud_mpm_step_fwd", "ud_mpm_step_bwd", "ud_mpm_sort_bins"} <= declared
    L = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/unidom_b200.h but not exported"
    L.ud_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.ud_version()


def test_workspace_query_and_validation_without_gpu(built_lib):
    from unidom_b200 import _lib
    L = _lib.lib()
    p = _lib.MpmParams()
    p.num_envs, p.n_particles, p.steps = 2, 1000, 16
    p.res = (ctypes.c_int32 * 3)(48, 32, 48)
    p.n_grid = 96
    p.dt, p.dx, p.inv_dx = 2e-4, 1 / 96, 96.0
    p.p_vol = (0.5 / 96) ** 2
    p.p_mass = p.p_vol
    p.gravity = (ctypes.c_double * 3)(0, -9.8, 0)
    p.n_primitive = 1
    fwd = L.ud_mpm_fwd_workspace_bytes(ctypes.byref(p))
    bwd = L.ud_mpm_bwd_workspace_bytes(ctypes.byref(p))
    assert 0 < fwd < bwd
    assert L.ud_mpm_num_keys(ctypes.byref(p)) == 12 * 8 * 12 * 64
    p.n_primitive = 9          # invalid
    assert L.ud_mpm_fwd_workspace_bytes(ctypes.byref(p)) == 0
    # p2g_mode: UD_P2G_ATOMIC / UD_P2G_DETERMINISTIC, optionally OR-ed with the UD_P2G_LIQUID_FAST flag; nothing else
    p.n_primitive = 1
    for mode, ok in ((0, True), (1, True), (_lib.UD_P2G_LIQUID_FAST, True), (1 | _lib.UD_P2G_LIQUID_FAST, True),
                     (2, False), (7, False), (2 | _lib.UD_P2G_LIQUID_FAST, False), (0x200, False)):
        p.p2g_mode = mode
        assert (L.ud_mpm_fwd_workspace_bytes(ctypes.byref(p)) > 0) == ok, mode
    p.p2g_mode = 0
    # null pointers are rejected before anything is enqueued
    rc = L.ud_mpm_step_fwd(ctypes.byref(p), None, None, None, None, None, None, 0, None)
    assert rc == -1


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "unidom_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
