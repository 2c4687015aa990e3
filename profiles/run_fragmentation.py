import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from unidom_b200 import confs
from unidom_b200.mpm_simulator import SimpleMPMSimulator
conf = confs.shape_elasto_plastic_conf()
sim = SimpleMPMSimulator(conf, 4, device='cuda')
state = bench.build_scene(sim, 3.9)
B, n = state.x.shape[:2]
action = torch.tensor([0.003, 0.0, 0.004, 0.0, 0.0, 0.0]).repeat(B, 1).cuda()
def runs(xord):
    base = (xord * np.float32(conf.inv_dx) - np.float32(0.5)).astype(np.int32)
    key = (base[:, 0] * 2048 + base[:, 1]) * 2048 + base[:, 2]
    nb = (len(key) + 127) // 128
    tot = 0; distinct = 0
    for b in range(nb):
        k = key[b*128:(b+1)*128]
        tot += 1 + int((k[1:] != k[:-1]).sum()); distinct += len(np.unique(k))
    return tot / nb, distinct / nb
with torch.no_grad():
    for it in range(10):
        base, key, perm = sim.sort_bins(state.x)
        p = perm[0].long().cpu().numpy()
        x0 = state.x[0].cpu().numpy()[p]
        new, _ = sim.step_jax(state, action)
        x1 = new.x[0].cpu().numpy()[p]
        print(it, 'runs/CTA at step start %.1f (distinct %.1f) ; at step end %.1f (distinct %.1f)' % (*runs(x0), *runs(x1)), 'ppc', n / len(np.unique(key[0].cpu().numpy())))
        state = new
