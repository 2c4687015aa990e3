"""CPU: the oracle against the reference's own ARTEFACTS (tests/golden/ref_artifacts.npz, copied from
/root/reference by tests/golden/gen_ref_artifacts.py): goal lattice, expert-demo gripper kinematics, whip_rope
primitive kinematics, reset's Lame parameters.  These are the only result-pinning files the reference ships."""
import os

import numpy as np
import torch

import util
from oracle import cloth as oc
from oracle import mpm as omp
from oracle import primitives as oP

ART = np.load(os.path.join(util.GOLD, "ref_artifacts.npz"))


def test_add_box_lattice_equals_shape_rope_goal():
    """goals/shape_rope/goal.npy is the particle cloud of shape_rope's reset (shape_rope_env.py:162-164):
    add_box(size=[0.25,0.006,0.006], init_pos=[0.5,0.01,0.5], density=3) at n_grid=128 (mpm_simulator.py:93-109)."""
    conf = omp.MPMConf(n_grid=128, res=(64, 6, 64), dt=0.5e-4, steps=133)
    x = omp.add_box(conf, [0.25, 0.006, 0.006], [0.5, 0.01, 0.5], density=3).numpy()
    goal = ART["shape_rope_goal"]
    assert x.shape == goal.shape == (582, 3)
    assert np.abs(x - goal).max() < 5e-5      # the goal was saved after the rope had rested for a few frames (|dz| <= 2e-5)


def test_cloth_gripper_trajectory_matches_expert_demo():
    """fold_cloth3/demo_0.pkl: after each env step the gripper sits at (place.x, 0.06, place.z): pins
    get_pnp_actions (3 approach + 10 lift + 20 move + 7 release), the /50 action scaling and the clip of robot_step."""
    acts, prim = torch.from_numpy(ART["cloth_actions"]), torch.from_numpy(ART["cloth_primitive0"])
    conf = oc.ClothConf()
    sim = oc.ClothSim(conf, oc.fold_cloth_mask(conf))
    st = sim.reset(1)
    st = st._replace(x=torch.from_numpy(ART["cloth_x0"]), primitive0=prim[0])
    with torch.no_grad():
        for t in range(2):
            sub = oc.get_pnp_actions(acts[t], st)
            assert sub.shape == (40, 1, 8)
            for a in sub:
                st = oc.step_batch(sim, st, a, 50)
            e = float((st.primitive0 - prim[t + 1]).abs().max())
            print(f"demo step {t}: gripper |err| {e:.2e}")
            assert e < 2e-6, (t, e)


def test_whip_rope_primitive_moves_49_50_of_command():
    """whip_rope/demo_0.pkl: position[0] advances by (S-1)/S of the commanded displacement per env step: the FK
    write to row S is dropped and copy_frame reads row S clamped to S-1 (primitives.py:185-194, mpm_simulator.py:365-373)."""
    S = int(ART["whip_prim_steps"])
    acts, pos = ART["whip_actions"][:, 0], ART["whip_prim_pos0"]
    p = oP.create_primitive(S, 0.0, 666.0, [0.5] * 3, [0.03, 0.03, 0.03], pos[0])
    for t in range(8):
        a = torch.from_numpy(acts[t]).clamp(-1, 1)
        p = oP.set_action(S, a, p)
        for f in range(S):
            p = oP.forward_kinematics(f, p)
        row0 = oP._row(p.position, S)                                     # copy_frame(S -> 0)
        p = p._replace(position=torch.cat([row0[None], p.position[1:]]))
        assert np.abs(row0.numpy() - pos[t + 1]).max() < 2e-6, t
        cmd = a[:3].numpy() * ART["whip_prim_action_scale"][0, :3]
        moved = pos[t + 1] - pos[t]
        nz = np.abs(cmd) > 1e-9
        assert np.allclose(moved[nz] / cmd[nz], (S - 1) / S, atol=1e-3)


def test_reset_lame_parameters_match_demo():
    """whip_rope demo state: mu = E/(2(1+nu)), lamda = E nu/((1+nu)(1-2nu)) for E=100, nu=0.1 (mpm_simulator.py:161-166)."""
    conf = omp.MPMConf(E=100.0, nu=0.1)
    st = omp.reset_state(conf, torch.zeros((4, 3)), [], 1)
    assert np.allclose([float(st.mu[0, 0]), float(st.lamda[0, 0])], ART["whip_mu_lamda"], rtol=1e-6)
