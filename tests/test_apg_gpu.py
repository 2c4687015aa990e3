"""The APG update kernels on the GPU: the one-kernel update (scrub + clip + mean over ranks + Adam, csrc/apg_fused.cu)
against the three-step path (ud_apg_scrub_clip -> all-reduce -> ud_adam_step) and against the host formula of
algorithms/apg/apg.py:233-240, 260-267 (tests/test_apg_cpu.py pins that formula's ordering on gloo)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ptr(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("scale,n", [(1e-3, 925964), (5.0, 4099), (1e-3, 7)])
def test_fused_update_world1_matches_three_step_path(built_lib, scale, n):
    """world = 1: no peer, the staged gradient is the rank's own.  scale 5.0 makes the norm exceed max_grad_norm (clip
    branch); NaN / inf entries exercise the scrub."""
    from unidom_b200 import _lib, apg
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    grad = (torch.randn(n, generator=g) * scale).to(dev)
    grad[1], grad[n // 2] = float("nan"), float("inf") if scale > 1 else 0.0
    params = torch.randn(n, generator=g).to(dev)
    opt = apg.Adam(n, 1e-3, dev)
    m2, v2 = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    stage = torch.zeros(2 * n, device=dev)
    flags = torch.zeros(64, dtype=torch.int32, device=dev)
    ps = torch.tensor([stage.data_ptr()], dtype=torch.int64, device=dev)
    pf = torch.tensor([flags.data_ptr()], dtype=torch.int64, device=dev)
    scratch = torch.zeros(8, device=dev)
    p_ref, p_fused = params.clone(), params.clone()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for t in range(1, 4):
        gt = grad * (1.0 + 0.1 * t)
        red, _ = apg.reduce_policy_gradient(gt, 0.3)
        p_ref = opt.step(p_ref, red)
        _lib.check(L.ud_apg_fused_update(_ptr(p_fused), _ptr(gt.contiguous()), _ptr(m2), _ptr(v2), n, 0.3, 1e-3, 0.9, 0.999,
                                         1e-8, t, 0, 1, _ptr(ps), _ptr(pf), _ptr(scratch), st), "ud_apg_fused_update")
        torch.cuda.synchronize()
        # the per-rank sum of squares is an fp32 atomic accumulation in both paths: equal up to its summation order
        err = float((p_fused - p_ref).abs().max())
        print(f"fused vs three-step update, n={n}, scale={scale}, t={t}: max |dp| {err:.3e}")
        assert err <= 2e-7 * float(p_ref.abs().max()), (t, err)
    assert torch.isfinite(p_fused).all()
    assert float((m2 - opt.m).abs().max()) <= 1e-6 * float(opt.m.abs().max() + 1e-30)


def test_fused_update_two_ranks_over_peer_memory(built_lib):
    """Two ranks under torchrun (NCCL + torch symmetric memory for the peer mapping): the fused kernel's replicas are
    bit-identical across ranks and agree with the NCCL path."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible devices (run with gpurun --gpus 2)")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531",
                        os.path.join(ROOT, "tests", "apg_fused_worker.py")], capture_output=True, text=True, env=env, timeout=300)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "fused update ok" in r.stdout
