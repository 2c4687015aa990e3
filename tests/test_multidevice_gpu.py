"""One process driving several devices, one host thread per device -- the contract `pmap` puts on the boundary (SURVEY
section 8b "Threading"; reference algorithms/apg/apg.py:271).  Kernel attributes (dynamic shared memory above 48 KB) are
per DEVICE, so a library that sets them once per process works on device 0 only: both adjoints are driven on device 1
from a second thread while device 0 runs, and each device's result must equal what the same device computes alone."""
import threading

import pytest
import torch

import util

pytestmark = pytest.mark.gpu


def _mpm_job(dev, out, key):
    from unidom_b200 import confs
    from unidom_b200.mpm_simulator import SimpleMPMSimulator
    torch.cuda.set_device(dev)
    conf = confs.shape_elasto_plastic_conf()
    conf.steps = 8
    B = 2
    sim = SimpleMPMSimulator(conf, B, device=dev, p2g_mode=1)          # deterministic P2G: bit-reproducible forward
    st = util.mini_plasticine(sim, B, seed=7)
    act = torch.tensor([[0.3, 0.0, 0.6, 0.0, 0.0, 0.1], [-0.2, 0.0, 0.4, 0.0, 0.0, 0.0]], device=dev)
    x = st.x.clone().requires_grad_(True)
    o, _ = sim.step_jax(st._replace(x=x), act)
    (gx,) = torch.autograd.grad((o.x * o.v).sum(), [x])
    torch.cuda.synchronize(dev)
    out[key] = (o.x.cpu(), o.F.cpu(), gx.cpu())


def _cloth_job(dev, out, key):
    from unidom_b200 import confs
    from unidom_b200.cloth_simulator import ClothSimulator
    torch.cuda.set_device(dev)
    conf = confs.ClothConf()
    B = 2
    sim = ClothSimulator(conf, B, None, confs.fold_cloth_mask(conf), device=dev)
    st = sim.reset_jax()
    act = torch.tensor([[0.3, 0.5, -0.2, 0.0, -0.1, 0.2, 0.4, 0.3], [0.6, -0.4, 0.1, 1.0, 0.0, 0.3, 0.0, 0.0]], device=dev)
    x = st.x.clone().requires_grad_(True)
    o, _ = sim.step_jax(st._replace(x=x), act)
    (gx,) = torch.autograd.grad((o.x * o.x).sum(), [x])
    torch.cuda.synchronize(dev)
    out[key] = (o.x.cpu(), gx.cpu())


@pytest.mark.parametrize("job", [_mpm_job, _cloth_job])
def test_two_devices_two_host_threads(built_lib, job):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible devices (run with gpurun --gpus 2)")
    d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
    alone = {}
    job(d1, alone, "d1")                      # device 1 first and alone: a per-process attribute flag would already fail here
    job(d0, alone, "d0")
    both = {}
    errs = []

    def run(dev, key):
        try:
            for _ in range(3):
                job(dev, both, key)
        except Exception as e:               # noqa: BLE001 -- surfaced below
            errs.append((key, repr(e)))
    ts = [threading.Thread(target=run, args=(d0, "d0")), threading.Thread(target=run, args=(d1, "d1"))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for key in ("d0", "d1"):
        for a, b in zip(alone[key], both[key]):
            if job is _mpm_job:
                assert torch.equal(a[..., :0], b[..., :0])      # shapes
                assert util.rel_err(a, b) < 1e-5, key           # the adjoint's fp32 REDs are order-dependent
            else:
                assert torch.equal(a, b), key                   # the cloth step is deterministic
    # the same scene on two devices gives the same forward
    assert torch.equal(alone["d0"][0], alone["d1"][0])
