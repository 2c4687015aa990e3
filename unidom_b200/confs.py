"""Scene configurations: the scalar fields of the reference's per-task DefaultConf classes.

Each class cites the reference conf it restates; `synthetic_*` helpers build the scaled-up scenes
BASELINE.json names.  Host-side only (no arithmetic of the hot path here).
"""
from dataclasses import dataclass, field
from typing import Tuple

from . import _lib


@dataclass
class MPMConf:
    n_grid: int = 64
    res: Tuple[int, int, int] = (32, 32, 32)
    dt: float = 1e-4
    steps: int = 16
    E: float = 100.0
    nu: float = 0.1
    ground_friction: float = 0.1
    gravity: Tuple[float, float, float] = (0.0, -9.8, 0.0)
    n_primitive: int = 1
    sdf_kind: int = _lib.UD_SDF_BOX
    use_position_control: bool = False
    p_rho: float = 1.0
    seed: int = 1
    task: str = ""

    @property
    def dx(self):
        return 1 / self.n_grid

    @property
    def inv_dx(self):
        return float(self.n_grid)

    @property
    def p_vol(self):
        return (self.dx * 0.5) ** 2

    @property
    def p_mass(self):
        return self.p_vol * self.p_rho


def shape_elasto_plastic_conf():
    """envs/shape_elasto_plastic.py:23-54 ("push_plasticine", BASELINE config 2)."""
    n_grid = 96
    return MPMConf(n_grid=n_grid, res=(n_grid // 2, n_grid // 3, n_grid // 2), dt=2e-4, steps=16, E=2, nu=0.2,
                   ground_friction=2, n_primitive=1, sdf_kind=_lib.UD_SDF_BOX, task="shape_elasto_plastic")


def shape_rope_conf():
    """envs/shape_rope_env.py:26-62."""
    n_grid = 128
    dt = 0.5e-4
    return MPMConf(n_grid=n_grid, res=(n_grid // 2, 6, n_grid // 2), dt=dt, steps=int(0.2 / 30 / dt), E=100, nu=0.1,
                   ground_friction=0.9, n_primitive=1, sdf_kind=_lib.UD_SDF_BOX, task="shape_rope")


def whip_rope_conf():
    """envs/whip_rope_env.py:27-73."""
    n_grid = 64
    dt = 1e-4
    return MPMConf(n_grid=n_grid, res=(n_grid // 2,) * 3, dt=dt, steps=int(0.007 / 1 / dt), E=100, nu=0.1,
                   ground_friction=0.1, n_primitive=1, sdf_kind=_lib.UD_SDF_BOX, use_position_control=True,
                   task="whip_rope")


def pour_water_conf(res=None):
    """envs/pour_water_env.py:28-60."""
    n_grid = 80
    dt = 3e-4
    return MPMConf(n_grid=n_grid, res=res or (n_grid // 3, n_grid // 4, n_grid // 3), dt=dt,
                   steps=int(0.007 / 1 / dt), E=0.00005, nu=0.4999, ground_friction=0.1, n_primitive=2,
                   sdf_kind=_lib.UD_SDF_CONTAINER, task="pour_water")


def build_shape_elasto_plastic(sim, density=3.0):
    """reset() of envs/shape_elasto_plastic.py:139-157.  density=3.9 gives the 50 625-particle
    synthetic scale-up of BASELINE config 2."""
    from .mpm_simulator import create_primitive
    conf = sim.conf
    state = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.2, 0.06, 0.12], init_pos=[0.5, 0.07, 0.5],
                        z_rotation_angle=0, material=2, density=density)
    state.primitives.append(create_primitive(conf, friction=0.1, softness=666, color=[0.5, 0.5, 0.5],
                                             size=[0.015, 0.06, 0.015], init_pos=[0.5, 0.01, 0.45]))
    return sim.reset_jax(state)


def build_shape_rope(sim, density=3):
    """reset() of envs/shape_rope_env.py:154-171 before its random pushes: a 0.25 x 0.006 x 0.006 plastic rope (582
    particles at density 3 = goals/shape_rope/goal.npy) and the box pusher."""
    from .mpm_simulator import create_primitive
    conf = sim.conf
    state = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.25, 0.006, 0.006], init_pos=[0.5, 0.01, 0.5],
                        z_rotation_angle=0, material=2, density=density)
    state.primitives.append(create_primitive(conf, friction=0.1, softness=666, color=[0.5, 0.5, 0.5],
                                             size=[0.015, 0.06, 0.015], init_pos=[0.5, 0.01, 0.45]))
    return sim.reset_jax(state)


def build_whip_rope(sim, density=2.75):
    """reset() of envs/whip_rope_env.py:119-137 before its random xz shift: a 0.38-long rope rotated by pi/2 about y and
    a 0.02 box gripper (position control) at [0.5, 0.01, 0.3]; 67 particles at the shipped density."""
    import math

    from .mpm_simulator import create_primitive
    conf = sim.conf
    state = sim.add_box(conf=conf, state=None, hardness=1.0, size=[0.38, 0.006, 0.006], init_pos=[0.5, 0.01, 0.5],
                        z_rotation_angle=math.pi / 2, material=1, density=density)
    state.primitives.append(create_primitive(conf, friction=0.1, softness=666, color=[0.5, 0.5, 0.5],
                                             size=[0.02, 0.02, 0.02], init_pos=[0.5, 0.01, 0.3]))
    return sim.reset_jax(state)


class ClothConf:
    """envs/fold_cloth3_env.py:18-37 (shared by fold_cloth1/3, unfold_cloth1/3; fold_cloth1_para randomises
    `stiffness` per env, fold_cloth1_para_env.py:15-33)."""
    N = 80
    gravity = 0.5
    stiffness = 900
    damping = 2
    dt = 2e-3
    max_v = 2.0
    small_num = 1e-8
    mu = 0.5
    seed = 1
    mem_saving_level = 2
    task = "fold_cloth3"

    @property
    def cell_size(self):
        return 1.0 / self.N

    @property
    def size(self):
        return int(self.N / 5.0)


class FoldCloth1ParaConf(ClothConf):
    """envs/fold_cloth1_para_env.py:15-33: the fold_cloth scene whose `stiffness` the env constructor overwrites
    (default 900; apg_para.py:326-339 draws one value per training iteration)."""
    stiffness = 9
    task = "fold_cloth1"
    use_substep_obs = True


class UnfoldClothConf(ClothConf):
    """envs/unfold_cloth3_env.py:17-35 (unfold_cloth1 alike): the fold_cloth scene with friction mu = 3, max_steps 15;
    reset = lattice + N(0, 1e-4^2) noise + 3 random pick-and-place folds (ClothEnv.random_fold)."""
    mu = 3
    task = "unfold_cloth3"


class FoldTshirtConf(ClothConf):
    """envs/fold_cloth_tshirt_env.py:20-38: N = 180, 3 573 nodes -> one thread-block cluster per env.
    Use with ClothEnv(conf, B, 5, tshirt_mask_from_image(conf, img), obs_stride=10)."""
    N = 180
    stiffness = 5000
    dt = 0.5e-3
    mu = 0.9
    task = "fold_tshirt"


def tshirt_mask_from_image(conf, img):
    """create_cloth_mask, fold_cloth_tshirt_env.py:52-71, from the already decoded + resized + rotated image: `img` is
    the (N/2, N/2, 3) uint8 array the reference gets from cv2.imread/resize/rotate of others/t-shirt.jpg (cv2 is not
    a dependency here); dark pixels (channel sum < 100) are cloth, centred on the N x N board."""
    import numpy as np
    size = conf.N // 2
    h = size // 2
    img = np.asarray(img)
    assert img.shape[:2] == (size, size), img.shape
    mask = (img.astype(np.int64).sum(-1) < 100).astype(np.float32)
    m = np.zeros((conf.N, conf.N), dtype=np.float32)
    m[conf.N // 2 - h:conf.N // 2 + h, conf.N // 2 - h:conf.N // 2 + h] = mask
    return m


def fold_cloth_mask(conf):
    """create_cloth_mask, envs/fold_cloth3_env.py:51-56: a 16 x 32 patch of the N x N board (512 nodes)."""
    import numpy as np
    N, size = conf.N, conf.size
    m = np.zeros((N, N), dtype=np.float32)
    m[size * 2:size * 3, size * 2:size * 4] = 1
    return m
