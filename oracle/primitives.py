"""Rigid primitives / colliders -- torch-CPU restatement (test infrastructure).

Follows core/engine/primitives/primitives.py, box.py, container.py.
All functions are per-environment (the reference vmaps them); tensors carry no
batch axis.  JAX indexing semantics are mirrored explicitly:
  * reads at row ``f+1 == S`` clamp to ``S-1``      (primitives.py:149,190)
  * writes at row ``f+1 == S`` are dropped          (primitives.py:186,192)
"""
from typing import NamedTuple

import torch

SDF_BOX = 0        # core/engine/primitives/box.py:6-18
SDF_CONTAINER = 1  # core/engine/primitives/container.py:8-16


class PrimitiveState(NamedTuple):  # primitives.py:9-23 (same field order)
    size: torch.Tensor
    dim: torch.Tensor
    friction: torch.Tensor
    softness: torch.Tensor
    color: torch.Tensor
    position: torch.Tensor
    rotation: torch.Tensor
    v: torch.Tensor
    w: torch.Tensor
    xyz_limit: torch.Tensor
    action_buffer: torch.Tensor
    action_scale: torch.Tensor
    min_dist: torch.Tensor
    dist_norm: torch.Tensor


def create_primitive(steps, friction, softness, color, size, init_pos, dtype=torch.float32):
    """primitives.py:31-60."""
    position = torch.zeros((steps, 3), dtype=dtype)
    position[0] = torch.as_tensor(init_pos, dtype=dtype)
    rotation = torch.tensor([[1.0, 0.0, 0.0, 0.0]], dtype=dtype).repeat(steps, 1)
    return PrimitiveState(
        size=torch.as_tensor(size, dtype=dtype),
        dim=torch.tensor([3], dtype=torch.int32),
        friction=torch.tensor(float(friction), dtype=dtype),
        softness=torch.tensor(float(softness), dtype=dtype),
        color=torch.as_tensor(color, dtype=dtype),
        position=position,
        rotation=rotation,
        v=torch.zeros((steps, 3), dtype=dtype),
        w=torch.zeros((steps, 3), dtype=dtype),
        xyz_limit=torch.tensor([[0.0, 1.0]] * 3, dtype=dtype),
        action_buffer=torch.zeros((6,), dtype=dtype),
        action_scale=torch.ones((6,), dtype=dtype),
        min_dist=torch.tensor(0, dtype=torch.int32),
        dist_norm=torch.tensor(0, dtype=torch.int32),
    )


def jclip(x, lo=None, hi=None):
    """jnp.clip = minimum(hi, maximum(lo, x)): a tie with a bound passes HALF the cotangent (lax.max/min JVP);
    torch.clamp would pass all of it."""
    if lo is not None:
        x = torch.maximum(x, torch.as_tensor(lo, dtype=x.dtype))
    if hi is not None:
        x = torch.minimum(x, torch.as_tensor(hi, dtype=x.dtype))
    return x


def length(x):
    """primitives.py:68-69: sqrt(x.x + 1e-12) over the last axis."""
    return torch.sqrt((x * x).sum(-1) + 1e-12)


def qmul(q, r):
    """primitives.py:73-81 (terms = outer(r, q))."""
    t = torch.outer(r, q)
    w = t[0, 0] - t[1, 1] - t[2, 2] - t[3, 3]
    x = t[0, 1] + t[1, 0] - t[2, 3] + t[3, 2]
    y = t[0, 2] + t[1, 3] + t[2, 0] - t[3, 1]
    z = t[0, 3] - t[1, 2] + t[2, 1] + t[3, 0]
    out = torch.stack([w, x, y, z])
    return out / jclip(torch.sqrt(out.dot(out)), 1e-12)


def w2quat(axis_angle):
    """primitives.py:84-92."""
    # jnp.linalg.norm = sqrt(sum(x*x)): NaN cotangent at exactly 0 (zero angular velocity), scrubbed later by
    # norm_grad's nan_to_num -- torch.linalg.norm would return a finite subgradient instead
    w = torch.sqrt((axis_angle * axis_angle).sum()) + 1e-12
    v = (axis_angle / w) * torch.sin(w / 2)
    return torch.cat([torch.cos(w / 2).reshape(1), v[:3]])


def qrot(rot, v):
    """primitives.py:95-103; v is (...,3)."""
    qvec = rot[1:4].expand_as(v)
    uv = torch.linalg.cross(qvec, v, dim=-1)
    uuv = torch.linalg.cross(qvec, uv, dim=-1)
    return v + 2 * (rot[0] * uv + uuv)


def inv_trans(pos, position, rotation):
    """primitives.py:106-110."""
    inv_quat = torch.stack([rotation[0], -rotation[1], -rotation[2], -rotation[3]])
    inv_quat = inv_quat / (torch.linalg.norm(inv_quat) + 1e-12)
    return qrot(inv_quat, pos - position)


def sdf_box(size, p):
    """box.py:6-18."""
    q = torch.abs(p) - size.reshape(3)
    qc = jclip(q, 0.0)
    out = length(qc)
    # NOTE box.py:11 clips q in place *before* the max -- the "inside" term
    # uses the clipped q, so tmp is max(clipped q) clipped to <= 0, i.e. 0.
    tmp = torch.where(qc[..., 1] > qc[..., 2], qc[..., 1], qc[..., 2])
    tmp = torch.where(qc[..., 0] > tmp, qc[..., 0], tmp)
    tmp = jclip(tmp, None, 0.0)
    return out + tmp


def sdf_container(size, p):
    """container.py:8-16 (cut hollow sphere, size = (r, h, t))."""
    r, h, t = size[0], size[1], size[2]
    w = torch.sqrt(r * r - h * h)
    q = torch.stack([length(p[..., [0, 2]]), p[..., 1]], dim=-1)
    mask = h * q[..., 0] < w * q[..., 1]
    val1 = length(q - torch.stack([w, h])) - t
    val2 = torch.abs(length(q) - r) - t
    return torch.where(mask, val1, val2)


_SDF = {SDF_BOX: sdf_box, SDF_CONTAINER: sdf_container}


def _row(table, f):
    """JAX gather semantics: out-of-range row index clamps."""
    return table[min(max(f, 0), table.shape[0] - 1)]


def sdf(f, grid_pos, state, kind):
    """primitives.py:112-114."""
    gp = inv_trans(grid_pos, _row(state.position, f), _row(state.rotation, f))
    return _SDF[kind](state.size, gp)


def _normal(grid_pos, state, kind):
    """primitives.py:117-136: central differences with d=1e-6 (in state dtype)."""
    d = 1.0e-6
    fn = _SDF[kind]
    comps = []
    for a in range(3):
        e = torch.zeros(3, dtype=grid_pos.dtype)
        e[a] = d
        inc = grid_pos + e
        dec = grid_pos + (-e)
        comps.append((0.5 / d) * (fn(state.size, inc) - fn(state.size, dec)))
    n = torch.stack(comps, dim=-1)
    return n / length(n)[..., None]


def normal(f, grid_pos, state, kind):
    """primitives.py:139-143."""
    gp = inv_trans(grid_pos, _row(state.position, f), _row(state.rotation, f))
    return qrot(_row(state.rotation, f), _normal(gp, state, kind))


def collider_v(f, grid_pos, dt, state):
    """primitives.py:146-153."""
    rot = _row(state.rotation, f)
    inv_quat = torch.stack([rot[0], -rot[1], -rot[2], -rot[3]])
    inv_quat = inv_quat / (torch.linalg.norm(inv_quat) + 1e-12)
    relative_pos = qrot(inv_quat, grid_pos - _row(state.position, f))
    new_pos = qrot(_row(state.rotation, f + 1), relative_pos) + _row(state.position, f + 1)
    return (new_pos - grid_pos) / dt


def collide(f, grid_pos, v_out, dt, state, kind):
    """primitives.py:156-182.  grid_pos, v_out: (...,3)."""
    dist = sdf(f, grid_pos, state, kind)
    influence = jclip(torch.exp(-dist * state.softness), None, 1.0)[..., None]
    D = normal(f, grid_pos, state, kind)
    cv = collider_v(f, grid_pos, dt, state)
    input_v = v_out - cv
    normal_component = (input_v * D).sum(-1, keepdim=True)
    grid_v_t = input_v - jclip(normal_component, None, 0.0) * D
    grid_v_t_norm = length(grid_v_t)[..., None]
    grid_v_t_friction = grid_v_t / grid_v_t_norm * torch.clamp(
        grid_v_t_norm + normal_component * state.friction, min=1e-12)
    grid_v_t_dot = (grid_v_t * grid_v_t).sum(-1, keepdim=True)
    flag = ((normal_component < 0) & (torch.sqrt(grid_v_t_dot) > 1e-12)).to(v_out.dtype).detach()
    grid_v_t = grid_v_t_friction * flag + grid_v_t * (1 - flag)
    return cv + input_v * (1 - influence) + grid_v_t * influence


def position_control(f, grid_pos, v_out, dt, state, kind):
    """primitives.py:232-239."""
    dist = sdf(f, grid_pos, state, kind)
    control_mask = dist < state.size[0] * 1.5
    return torch.where(control_mask[..., None], (_row(state.v, f) / dt).reshape(1, 1, 1, 3), v_out)


def forward_kinematics(f, state):
    """primitives.py:185-194.  Write to row f+1 is dropped when f+1 == S."""
    S = state.position.shape[0]
    position, rotation = state.position, state.rotation
    if f + 1 < S:
        new_p = state.position[f] + state.v[f]
        position = torch.cat([position[:f + 1], new_p[None], position[f + 2:]], dim=0)
        new_r = qmul(w2quat(state.w[f]), state.rotation[f])
        rotation = torch.cat([rotation[:f + 1], new_r[None], rotation[f + 2:]], dim=0)
    position = jclip(position, -2, 2)
    return state._replace(position=position, rotation=rotation)


def set_action(n_substeps, action, state):
    """primitives.py:212-229 (set_velocity fills rows 0..n_substeps-1)."""
    vrow = action[:3] * state.action_scale[:3] / n_substeps
    wrow = action[3:] * state.action_scale[3:] / n_substeps
    S = state.v.shape[0]
    n = min(n_substeps, S)
    v = torch.cat([vrow[None].expand(n, 3), state.v[n:]], dim=0)
    w = torch.cat([wrow[None].expand(n, 3), state.w[n:]], dim=0)
    return state._replace(action_buffer=action, v=v, w=w)
