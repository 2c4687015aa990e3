// Grid-side kernels of the MLS-MPM step: primitive forward kinematics, the fused grid update
// (normalise + gravity + colliders + ground friction + boundary) and their adjoints.
// sm_100a; compile this file with --fmad=false (see mpm_cell.cuh).
// Reference: mpm_simulator.py:277-313,365-373,413-429; primitives.py:73-92,185-229.
#include "mpm_cell.cuh"
#include "mpm_internal.h"

namespace ud {

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// primitives.py:73-81 (terms = outer(r, q)), normalised
template <class T> UD_DEV void qmul(const T q[4], const T r[4], T out[4]) {
  T w = r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3];
  T x = r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2];
  T y = r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1];
  T z = r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0];
  T nrm = s_max_c(s_sqrt(w * w + x * x + y * y + z * z), 1e-12f);
  out[0] = w / nrm;
  out[1] = x / nrm;
  out[2] = y / nrm;
  out[3] = z / nrm;
}
// primitives.py:84-92
template <class T> UD_DEV void w2quat(const T aa[3], T out[4]) {
  T w = s_sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]) + 1e-12f;
  T sn = s_sin(w / 2.f);
  out[0] = s_cos(w / 2.f);
  out[1] = (aa[0] / w) * sn;
  out[2] = (aa[1] / w) * sn;
  out[3] = (aa[2] / w) * sn;
}

UD_DEV float clipf(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }

// ------------------------------------------------------------------------------------------------
// step prologue for the primitives (mpm_simulator.py:419-423 + the FK of every substep, which only
// depends on the action): action clip, set_action, rows 1..S-1 of position/rotation, copy_frame.
// One thread per (env, primitive); S is 16..133.
// Row f+1 == S is never written and reads of it clamp to S-1 (JAX scatter drop / gather clamp).
// ------------------------------------------------------------------------------------------------
__global__ void k_fk_fwd(MpmConst k, ud_mpm_state in, const float* __restrict__ action, ud_mpm_state out,
                         float* __restrict__ fk_pos, float* __restrict__ fk_rot, float* __restrict__ fk_vw,
                         float* __restrict__ fk_act, int write_out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k.B * k.n_prim) return;
  int env = t / k.n_prim, q = t % k.n_prim;
  const ud_primitive& pi = in.prim[q];
  const ud_primitive& po = out.prim[q];
  const int S = k.S;
  float a[6], vw[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    a[j] = clipf(action[(size_t)env * 6 * k.n_prim + 6 * q + j], -1.f, 1.f);
    vw[j] = a[j] * pi.action_scale[env * 6 + j] / (float)S;
    fk_act[(size_t)t * 6 + j] = a[j];
    fk_vw[(size_t)t * 6 + j] = vw[j];
  }
  float* tp = fk_pos + (size_t)t * (S + 1) * 3;
  float* tr = fk_rot + (size_t)t * (S + 1) * 4;
  float pos[3], rot[4], dq[4];
#pragma unroll
  for (int j = 0; j < 3; ++j) pos[j] = pi.position[((size_t)env * S) * 3 + j];
#pragma unroll
  for (int j = 0; j < 4; ++j) rot[j] = pi.rotation[((size_t)env * S) * 4 + j];
  w2quat(vw + 3, dq);
  for (int f = 0; f < S; ++f) {
#pragma unroll
    for (int j = 0; j < 3; ++j) tp[f * 3 + j] = pos[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) tr[f * 4 + j] = rot[j];
    if (f + 1 < S) {
#pragma unroll
      for (int j = 0; j < 3; ++j) pos[j] = clipf(pos[j] + vw[j], -2.f, 2.f);
      float nr[4];
      qmul(dq, rot, nr);
#pragma unroll
      for (int j = 0; j < 4; ++j) rot[j] = nr[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) tp[S * 3 + j] = pos[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) tr[S * 4 + j] = rot[j];
  if (!write_out) return;
  // output leaves: rows 1..S-1 as computed, row 0 = copy_frame(S -> 0) = row S-1
  for (int f = 0; f < S; ++f) {
    int src = f == 0 ? S - 1 : f;
#pragma unroll
    for (int j = 0; j < 3; ++j) po.position[((size_t)env * S + f) * 3 + j] = tp[src * 3 + j];
#pragma unroll
    for (int j = 0; j < 4; ++j) po.rotation[((size_t)env * S + f) * 4 + j] = tr[src * 4 + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      po.v[((size_t)env * S + f) * 3 + j] = vw[j];
      po.w[((size_t)env * S + f) * 3 + j] = vw[3 + j];
    }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    po.action_buffer[env * 6 + j] = a[j];
    po.action_scale[env * 6 + j] = pi.action_scale[env * 6 + j];
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) po.size[env * 3 + j] = pi.size[env * 3 + j];
  po.friction[env] = pi.friction[env];
  po.softness[env] = pi.softness[env];
}

void launch_fk_fwd(const MpmConst& k, const ud_mpm_state* in, const float* action, ud_mpm_state* out,
                   const MpmWs& ws, cudaStream_t st) {
  int n = k.B * k.n_prim;
  if (n == 0) return;
  k_fk_fwd<<<cdiv(n, 64), 64, 0, st>>>(k, *in, action, out ? *out : *in, ws.fk_pos, ws.fk_rot, ws.fk_vw,
                                       ws.fk_act, out != nullptr);
}

// ------------------------------------------------------------------------------------------------
// Fused grid update, one thread per cell, dense over B*G.  Cells that received no mass are skipped:
// nothing gathers them with a non-zero weight (the reference computes gravity/collider values there
// that are never read).
// ------------------------------------------------------------------------------------------------
template <class T>
UD_DEV void load_prim_f(const MpmConst& k, const ud_mpm_state& in, const float* fk_pos, const float* fk_rot,
                        const float* fk_vw, int env, int q, int f, PrimIn<float>& pr) {
  size_t t = (size_t)env * k.n_prim + q;
  const float* tp = fk_pos + t * (k.S + 1) * 3;
  const float* tr = fk_rot + t * (k.S + 1) * 4;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    pr.pos_f[j] = tp[f * 3 + j];
    pr.pos_f1[j] = tp[(f + 1) * 3 + j];
    pr.size[j] = in.prim[q].size[env * 3 + j];
    pr.v_f[j] = fk_vw[t * 6 + j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    pr.rot_f[j] = tr[f * 4 + j];
    pr.rot_f1[j] = tr[(f + 1) * 4 + j];
  }
  pr.friction = in.prim[q].friction[env];
  pr.softness = in.prim[q].softness[env];
}

__global__ void __launch_bounds__(128)
k_grid_fwd(MpmConst k, const float4* grid_in, float4* grid_out, int f,
           ud_mpm_state in, const float* __restrict__ fk_pos, const float* __restrict__ fk_rot,
           const float* __restrict__ fk_vw) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)k.B * k.G) return;
  float4 g = grid_in[idx];
  if (!(g.w > 0.f)) {
    if (grid_out != grid_in) grid_out[idx] = g;
    return;
  }
  int env = (int)(idx / k.G);
  int c = (int)(idx - (size_t)env * k.G);
  int ck = c % k.rz, cj = (c / k.rz) % k.ry, ci = c / (k.rz * k.ry);
  PrimIn<float> prims[UD_MAX_PRIM];
  for (int q = 0; q < k.n_prim; ++q) load_prim_f<float>(k, in, fk_pos, fk_rot, fk_vw, env, q, f, prims[q]);
  float p[3] = {g.x, g.y, g.z}, v[3];
  cell_update<float>(k, ci, cj, ck, p, g.w, in.friction[env], prims, v);
  grid_out[idx] = make_float4(v[0], v[1], v[2], g.w);
}

void launch_grid_fwd(const MpmConst& k, const float4* grid_in, float4* grid_out, int substep,
                     const ud_mpm_state* in, const MpmWs& ws, cudaStream_t st) {
  k_grid_fwd<<<cdiv((long long)k.B * k.G, 128), 128, 0, st>>>(k, grid_in, grid_out, substep, *in, ws.fk_pos,
                                                              ws.fk_rot, ws.fk_vw);
}

}  // namespace ud
