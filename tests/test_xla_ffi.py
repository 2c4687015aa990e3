"""csrc/xla_ffi.cc (the jax.ffi handlers) compiled against the test-only stand-in of XLA's FFI API (tests/xla_stub) and
CALLED through it: what the adapter is responsible for -- buffer order, attribute decoding, workspace plumbing, error
mapping -- is checked against direct C-ABI calls.  ABI compatibility with a real XLA is NOT (no jax in this image)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HANDLERS = ("ud_xla_mpm_step_fwd", "ud_xla_mpm_step_bwd", "ud_xla_cloth_step_fwd", "ud_xla_cloth_step_bwd",
            "ud_xla_mpm_step_fwd_taped", "ud_xla_mpm_step_bwd_taped", "ud_xla_cloth_multi_step_fwd",
            "ud_xla_cloth_multi_step_bwd", "ud_xla_chamfer_fwd", "ud_xla_chamfer_bwd", "ud_xla_l2_fwd", "ud_xla_l2_bwd",
            "ud_xla_apg_scrub_clip", "ud_xla_adam_step")


@pytest.fixture(scope="module")
def xla(built_lib):
    from unidom_b200 import build
    path, kind = build.build_xla_adapter()
    if kind != "stub":
        pytest.skip("jax is installed: the adapter was built against the real XLA headers, the stub driver is absent")
    return C.CDLL(path)


def call(lib, name, stream, args, rets, attrs=()):
    """args / rets: lists of (data pointer, dims, element bytes)."""
    bufs = list(args) + list(rets)
    ptrs = (C.c_void_p * len(bufs))(*[b[0] for b in bufs])
    ranks = (C.c_int * len(bufs))(*[len(b[1]) for b in bufs])
    flat = [d for b in bufs for d in b[1]]
    dims = (C.c_int64 * max(len(flat), 1))(*flat)
    eb = (C.c_int * len(bufs))(*[b[2] for b in bufs])
    names = (C.c_char_p * max(len(attrs), 1))(*[a[0].encode() for a in attrs])
    isd = (C.c_int * max(len(attrs), 1))(*[int(isinstance(a[1], float)) for a in attrs])
    ai = (C.c_int64 * max(len(attrs), 1))(*[0 if isinstance(a[1], float) else int(a[1]) for a in attrs])
    ad = (C.c_double * max(len(attrs), 1))(*[float(a[1]) if isinstance(a[1], float) else 0.0 for a in attrs])
    err = C.create_string_buffer(512)
    rc = lib.ud_stub_call(getattr(lib, name), C.c_void_p(stream), len(args), len(rets), ptrs, ranks, dims, eb, len(attrs),
                          names, isd, ai, ad, err, 512)
    return rc, err.value.decode()


def test_adapter_compiles_exports_every_handler_and_maps_errors(xla):
    for h in HANDLERS:
        assert hasattr(xla, h), h
    # null buffers: the C ABI rejects them, the adapter turns the status into an FFI error that carries ud_last_error()
    rc, msg = call(xla, "ud_xla_l2_fwd", 0, [(0, (2, 5, 3), 4), (0, (5, 3), 4)], [(0, (2,), 4)])
    assert rc != 0 and "ud_l2_fwd" in msg, (rc, msg)
    rc, msg = call(xla, "ud_xla_adam_step", 0, [(0, (4,), 4)] * 4, [(0, (4,), 4)] * 3, [("world_size", 1), ("lr", 1e-3)])
    assert rc != 0 and "attribute" in msg, (rc, msg)          # b1, b2, eps, t missing


@pytest.mark.gpu
def test_handlers_match_direct_calls_on_the_gpu(xla):
    import util
    from unidom_b200 import _lib, confs
    from unidom_b200.mpm_simulator import SimpleMPMSimulator, flatten_state
    dev = torch.device("cuda", 0)
    st_ptr = torch.cuda.current_stream(dev).cuda_stream
    # ---- l2 reward
    g = torch.Generator().manual_seed(0)
    x = torch.rand((3, 50, 3), generator=g).to(dev)
    y = torch.rand((50, 3), generator=g).to(dev)
    out = torch.empty(3, device=dev)
    rc, msg = call(xla, "ud_xla_l2_fwd", st_ptr, [(x.data_ptr(), x.shape, 4), (y.data_ptr(), y.shape, 4)], [(out.data_ptr(), (3,), 4)])
    assert rc == 0, msg
    ref = torch.sqrt(((x - y[None]) ** 2).mean(-1)).mean(1)
    assert util.rel_err(out, ref) < 1e-6
    # ---- MPM forward through the handler vs the simulator (deterministic P2G: bit-identical)
    conf = confs.shape_elasto_plastic_conf()
    conf.steps = 4
    B = 2
    sim = SimpleMPMSimulator(conf, B, device=dev, p2g_mode=_lib.UD_P2G_DETERMINISTIC)
    state = util.mini_plasticine(sim, B, seed=2)
    act = torch.tensor([[0.2, 0.0, 0.5, 0.0, 0.0, 0.1], [-0.3, 0.0, 0.4, 0.0, 0.0, 0.0]], device=dev)
    with torch.no_grad():
        ref_state, _ = sim.step_jax(state, act)
    L = _lib.lib()
    p = sim.params(B=B, n=state.x.shape[1])
    ws = torch.empty(L.ud_mpm_fwd_workspace_bytes(C.byref(p)) + 256, dtype=torch.uint8, device=dev)
    ws_al = ws[(-ws.data_ptr()) % 256:][:ws.numel() - 256]
    prim = state.primitives[0]
    # the adapter's buffer order: material, h, action, then ud_mpm_state order: x v C F J friction mu lamda, then per
    # primitive size friction softness position rotation v w action_buffer action_scale
    leaves = [state.x, state.v, state.C, state.F, state.J, state.friction, state.mu, state.lamda, prim.size, prim.friction,
              prim.softness, prim.position, prim.rotation, prim.v, prim.w, prim.action_buffer, prim.action_scale]
    leaves = [t.to(torch.float32).contiguous() for t in leaves]
    outs = [torch.empty_like(t) for t in leaves]
    args = [(sim._material_dev.data_ptr(), sim._material_dev.shape, 4), (sim._h_dev.data_ptr(), sim._h_dev.shape, 4),
            (act.data_ptr(), act.shape, 4)] + [(t.data_ptr(), t.shape, 4) for t in leaves]
    rets = [(t.data_ptr(), t.shape, 4) for t in outs] + [(ws_al.data_ptr(), (ws_al.numel(),), 1)]
    attrs = [("steps", conf.steps), ("res_x", conf.res[0]), ("res_y", conf.res[1]), ("res_z", conf.res[2]),
             ("n_grid", conf.n_grid), ("dt", float(conf.dt)), ("p_rho", float(conf.p_rho)), ("gravity_x", float(conf.gravity[0])),
             ("gravity_y", float(conf.gravity[1])), ("gravity_z", float(conf.gravity[2])), ("n_primitive", 1), ("sdf_kind", 0),
             ("use_position_control", 0), ("p2g_mode", 1)]
    rc, msg = call(xla, "ud_xla_mpm_step_fwd", st_ptr, args, rets, attrs)
    assert rc == 0, msg
    torch.cuda.synchronize()
    for k, t in zip(("x", "v", "C", "F", "J"), outs[:5]):
        assert torch.equal(t, getattr(ref_state, k)), k
    assert torch.equal(outs[11], ref_state.primitives[0].position)
    # a workspace that is one byte short must come back as an FFI error naming the entry point
    rets[-1] = (ws_al.data_ptr(), (ws_al.numel() - 4096,), 1)
    rc, msg = call(xla, "ud_xla_mpm_step_fwd", st_ptr, args, rets, attrs)
    assert rc != 0 and "workspace" in msg, (rc, msg)
