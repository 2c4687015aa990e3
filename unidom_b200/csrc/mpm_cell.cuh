// Per-cell grid update of the MLS-MPM substep, written ONCE as a template over the scalar type:
//   T = float    -> forward kernel
//   T = Dual<N>  -> forward-mode tangents, contracted with the incoming cotangent in the adjoint
//                   kernel (cells are few compared with particles, so N-wide forward mode is cheap
//                   and cannot drift from the forward arithmetic, FD normals included).
// Reference: mpm_simulator.py:283-313 (normalise, gravity, colliders, ground friction, boundary),
// primitives.py:95-182,232-239 (quaternion utils, sdf, FD normal, collider velocity, collide,
// position control), box.py:6-18, container.py:8-16.
// This header must be compiled with --fmad=false: the FD normals (d = 1e-6 in fp32) are
// sensitive to contraction.
#pragma once
#include "mpm_particle.cuh"

namespace ud {

template <int N>
struct Dual {
  float v;
  float d[N];
};

// ---- scalar ops, float flavour
UD_DEV float s_val(float a) { return a; }
UD_DEV float s_sqrt(float a) { return sqrtf(a); }
UD_DEV float s_exp(float a) { return expf(a); }
UD_DEV float s_abs(float a) { return fabsf(a); }
UD_DEV float s_sin(float a) { return sinf(a); }
UD_DEV float s_cos(float a) { return cosf(a); }
UD_DEV float s_const(float, float c) { return c; }  // constant of the same type as the 1st arg
UD_DEV float s_sel(bool c, float a, float b) { return c ? a : b; }

// ---- scalar ops, dual flavour
template <int N> UD_DEV float s_val(const Dual<N>& a) { return a.v; }
template <int N> UD_DEV Dual<N> s_const(const Dual<N>&, float c) {
  Dual<N> r; r.v = c;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = 0.f;
  return r;
}
template <int N> UD_DEV Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N> UD_DEV Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N> UD_DEV Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r; r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N> UD_DEV Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N> UD_DEV Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = a.v / b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
template <int N> UD_DEV Dual<N> operator+(const Dual<N>& a, float b) { Dual<N> r = a; r.v += b; return r; }
template <int N> UD_DEV Dual<N> operator+(float b, const Dual<N>& a) { Dual<N> r = a; r.v += b; return r; }
template <int N> UD_DEV Dual<N> operator-(const Dual<N>& a, float b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> UD_DEV Dual<N> operator-(float b, const Dual<N>& a) { Dual<N> r = -a; r.v += b; return r; }
template <int N> UD_DEV Dual<N> operator*(const Dual<N>& a, float b) {
  Dual<N> r; r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
  return r;
}
template <int N> UD_DEV Dual<N> operator*(float b, const Dual<N>& a) { return a * b; }
template <int N> UD_DEV Dual<N> operator/(const Dual<N>& a, float b) {
  Dual<N> r; r.v = a.v / b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b;
  return r;
}
template <int N> UD_DEV Dual<N> s_sqrt(const Dual<N>& a) {
  Dual<N> r; r.v = sqrtf(a.v);
  float g = 0.5f / r.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * g;
  return r;
}
template <int N> UD_DEV Dual<N> s_exp(const Dual<N>& a) {
  Dual<N> r; r.v = expf(a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * r.v;
  return r;
}
template <int N> UD_DEV Dual<N> s_sin(const Dual<N>& a) {
  Dual<N> r; r.v = sinf(a.v); float c = cosf(a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c;
  return r;
}
template <int N> UD_DEV Dual<N> s_cos(const Dual<N>& a) {
  Dual<N> r; r.v = cosf(a.v); float c = -sinf(a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c;
  return r;
}
template <int N> UD_DEV Dual<N> s_abs(const Dual<N>& a) {
  // jnp.abs gradient = sign(x) (0 at 0)
  Dual<N> r; r.v = fabsf(a.v);
  float g = a.v > 0.f ? 1.f : (a.v < 0.f ? -1.f : 0.f);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * g;
  return r;
}
template <int N> UD_DEV Dual<N> s_sel(bool c, const Dual<N>& a, const Dual<N>& b) { return c ? a : b; }

// clip helpers (gradient 1 inside the closed range, 0 outside -- torch.clamp / jnp.clip away from ties)
template <class T> UD_DEV T s_max_c(const T& a, float lo) { return s_sel(s_val(a) >= lo, a, s_const(a, lo)); }
template <class T> UD_DEV T s_min_c(const T& a, float hi) { return s_sel(s_val(a) <= hi, a, s_const(a, hi)); }

// ---- small vector helpers
template <class T> UD_DEV T dot3(const T a[3], const T b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> UD_DEV void cross3(const T a[3], const T b[3], T r[3]) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
// primitives.py:68-69
template <class T> UD_DEV T length3(const T a[3]) { return s_sqrt(dot3(a, a) + 1e-12f); }
template <class T> UD_DEV T length2(const T& a, const T& b) { return s_sqrt(a * a + b * b + 1e-12f); }

// primitives.py:95-103
template <class T> UD_DEV void qrot(const T rot[4], const T v[3], T out[3]) {
  T qv[3] = {rot[1], rot[2], rot[3]};
  T uv[3], uuv[3];
  cross3(qv, v, uv);
  cross3(qv, uv, uuv);
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = v[i] + 2.f * (rot[0] * uv[i] + uuv[i]);
}
// inverse quaternion as in primitives.py:108-109 / :147-148
template <class T> UD_DEV void inv_quat(const T rot[4], T iq[4]) {
  T nrm = s_sqrt(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2] + rot[3] * rot[3]) + 1e-12f;
  iq[0] = rot[0] / nrm;
  iq[1] = (-rot[1]) / nrm;
  iq[2] = (-rot[2]) / nrm;
  iq[3] = (-rot[3]) / nrm;
}

// box.py:6-18 (q is clipped before the max, so the "inside" term is identically 0)
template <class T> UD_DEV T sdf_box(const T size[3], const T p[3]) {
  T q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) q[i] = s_max_c(s_abs(p[i]) - size[i], 0.f);
  T out = length3(q);
  T tmp = s_sel(s_val(q[1]) > s_val(q[2]), q[1], q[2]);
  tmp = s_sel(s_val(q[0]) > s_val(tmp), q[0], tmp);
  tmp = s_min_c(tmp, 0.f);
  return out + tmp;
}
// container.py:8-16
template <class T> UD_DEV T sdf_container(const T size[3], const T p[3]) {
  const T &r = size[0], &h = size[1], &t = size[2];
  T w = s_sqrt(r * r - h * h);
  T q0 = length2(p[0], p[2]);
  T q1 = p[1];
  bool mask = s_val(h * q0) < s_val(w * q1);
  T val1 = length2(q0 - w, q1 - h) - t;
  T val2 = s_abs(length2(q0, q1) - r) - t;
  return s_sel(mask, val1, val2);
}
template <class T> UD_DEV T sdf_local(int kind, const T size[3], const T p[3]) {
  return kind == 0 ? sdf_box(size, p) : sdf_container(size, p);
}

template <class T>
struct PrimIn {
  T pos_f[3], rot_f[4], pos_f1[3], rot_f1[4], size[3], friction, v_f[3];
  float softness;
};
constexpr int PRIM_NIN = 21;  // differentiable scalars per primitive, in the order above

// primitives.py:156-182
template <class T>
UD_DEV void collide_cell(int kind, float dt, const float gpos[3], const PrimIn<T>& pr, T v[3]) {
  T iq[4];
  inv_quat(pr.rot_f, iq);
  T rel[3], gp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) rel[i] = gpos[i] - pr.pos_f[i];
  qrot(iq, rel, gp);
  T dist = sdf_local(kind, pr.size, gp);
  T influence = s_min_c(s_exp((-dist) * pr.softness), 1.f);
  // FD normal, primitives.py:117-136
  const float d = 1.e-6f;
  T nl[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    T inc[3] = {gp[0], gp[1], gp[2]}, dec[3] = {gp[0], gp[1], gp[2]};
    inc[a] = inc[a] + d;
    dec[a] = dec[a] + (-d);
    nl[a] = (0.5f / d) * (sdf_local(kind, pr.size, inc) - sdf_local(kind, pr.size, dec));
  }
  T nlen = length3(nl);
#pragma unroll
  for (int a = 0; a < 3; ++a) nl[a] = nl[a] / nlen;
  T D[3];
  qrot(pr.rot_f, nl, D);
  // collider velocity, primitives.py:146-153 (relative_pos == gp)
  T np[3], cv[3];
  qrot(pr.rot_f1, gp, np);
#pragma unroll
  for (int i = 0; i < 3; ++i) cv[i] = ((np[i] + pr.pos_f1[i]) - gpos[i]) / dt;
  T iv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) iv[i] = v[i] - cv[i];
  T nc = dot3(iv, D);
  T ncm = s_min_c(nc, 0.f);
  T vt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) vt[i] = iv[i] - ncm * D[i];
  T vtn = length3(vt);
  T fr = s_max_c(vtn + nc * pr.friction, 1e-12f);
  bool flag = s_val(nc) < 0.f && sqrtf(s_val(dot3(vt, vt))) > 1e-12f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T vtf = vt[i] / vtn * fr;
    T vsel = s_sel(flag, vtf, vt[i]);
    v[i] = cv[i] + iv[i] * (1.f - influence) + vsel * influence;
  }
}

// primitives.py:232-239
template <class T>
UD_DEV void position_control_cell(int kind, float dt, const float gpos[3], const PrimIn<T>& pr, T v[3]) {
  T iq[4];
  inv_quat(pr.rot_f, iq);
  T rel[3], gp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) rel[i] = gpos[i] - pr.pos_f[i];
  qrot(iq, rel, gp);
  T dist = sdf_local(kind, pr.size, gp);
  bool mask = s_val(dist) < s_val(pr.size[0] * 1.5f);
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = s_sel(mask, pr.v_f[i] / dt, v[i]);
}

// ground friction (mpm_simulator.py:297-307) and boundary (:310-313) for one cell, in place
template <class T>
UD_DEV void cell_ground_boundary(const MpmConst& k, int ci, int cj, int ck, const T& sfric, T v[3]) {
  const float ie[3] = {(float)ci * 1e-30f, (float)cj * 1e-30f, (float)ck * 1e-30f};
  if (cj < 3 && s_val(v[1]) <= 0.f) {
    T lin = v[1] + 1e-30f;
    T vit[3] = {v[0] - ie[0], (v[1] - lin) - ie[1], v[2] - ie[2]};
    T a[3] = {vit[0] + 1e-12f, vit[1] + 1e-12f, vit[2] + 1e-12f};
    T lit = s_sqrt(dot3(a, a));
    T sc = s_max_c(1.f + sfric * lin / lit, 0.f);
    v[0] = sc * (vit[0] + ie[0]);
    v[1] = s_const(v[1], 0.f);
    v[2] = sc * (vit[2] + ie[2]);
  }
  // upper wall tests n_grid, not res
  const int cidx[3] = {ci, cj, ck};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    bool cond = (cidx[d] < 3 && s_val(v[d]) < 0.f) || (cidx[d] > k.n_grid - 3 && s_val(v[d]) > 0.f);
    if (cond) v[d] = s_const(v[d], 0.f);
  }
}

// mpm_simulator.py:283-313 for one cell with m > 0.  p = scattered momentum, m = scattered mass.
// `prim_of(q, pr)` fills primitive q and returns false when the primitive is to be skipped (the
// adjoint skips primitives whose influence on this cell is below 1e-12; the forward never skips).
template <class T, class PF>
UD_DEV void cell_update(const MpmConst& k, int ci, int cj, int ck, const T p[3], const T& m, const T& sfric,
                        PF prim_of, T v[3]) {
  // mpm_simulator.py:283-285: where(m > 0, p / m, p) + dt * gravity (empty cells keep p = 0 and get dt * gravity)
  const bool has_mass = s_val(m) > 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = (has_mass ? p[i] / m : p[i]) + k.gdt[i];
  const float gpos[3] = {(float)ci * k.dx, (float)cj * k.dx, (float)ck * k.dx};
  for (int q = 0; q < k.n_prim; ++q) {
    PrimIn<T> pr;
    if (!prim_of(q, pr)) continue;
    if (k.pos_control)
      position_control_cell(k.sdf_kind, k.dt, gpos, pr, v);
    else
      collide_cell(k.sdf_kind, k.dt, gpos, pr, v);
  }
  cell_ground_boundary(k, ci, cj, ck, sfric, v);
}

// Cells of the boundary shell are the only EMPTY cells a particle can gather with a non-zero weight (a particle
// outside the grid gathers with clamped indices); the reference updates every cell, the kernels
// update cells with mass plus the empty cells of this shell.
// (Only the outermost layer: a particle more than 1.5 cells outside the low faces would wrap to layers r-2, r-3
// under JAX's negative-index rule; that far out the simulation is void anyway.)
UD_DEV bool cell_in_shell(const MpmConst& k, int ci, int cj, int ck) {
  return ci == 0 || cj == 0 || ck == 0 || ci == k.rx - 1 || cj == k.ry - 1 || ck == k.rz - 1;
}

// Does primitive `pr` act on the cell at gpos?  collide: influence = min(exp(-dist*softness),1) >= 1e-12
// (below that every derivative w.r.t. the primitive is < 1e-12 relative and d v_out / d v_in = I);
// position control: the mask itself.
UD_DEV bool prim_active(const MpmConst& k, const float gpos[3], const PrimIn<float>& pr) {
  float iq[4], rel[3], gp[3];
  inv_quat(pr.rot_f, iq);
  for (int i = 0; i < 3; ++i) rel[i] = gpos[i] - pr.pos_f[i];
  qrot(iq, rel, gp);
  float dist = sdf_local(k.sdf_kind, pr.size, gp);
  if (k.pos_control) return dist < pr.size[0] * 1.5f;
  return fminf(expf(-dist * pr.softness), 1.f) >= 1e-12f;
}

// ================================================================================================
// Hand-written reverse mode of collide_cell / position_control_cell (float only).  Used by the grid
// adjoint for cells a primitive acts on; checked against the Dual<N> Jacobian of the templates above
// (tests/test_hostmath_cpu.py) and against torch autograd of the oracle (tests/test_mpm_gpu.py).
// Conventions: every *_bwd ACCUMULATES into its output cotangents.
// PrimGrad layout = the PRIM_NIN order: pos_f(3) rot_f(4) pos_f1(3) rot_f1(4) size(3) friction v_f(3).
// ================================================================================================
struct PrimGrad {
  float g[PRIM_NIN];
};

UD_DEV void cross3f(const float a[3], const float b[3], float r[3]) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}

// out = v + 2 (r0 (q x v) + q x (q x v))
UD_DEV void qrot_bwd(const float rot[4], const float v[3], const float gout[3], float grot[4], float gv[3]) {
  const float q[3] = {rot[1], rot[2], rot[3]};
  float uv[3];
  cross3f(q, v, uv);
  float guv[3], guuv[3], t[3];
  for (int i = 0; i < 3; ++i) {
    gv[i] += gout[i];
    guv[i] = 2.f * rot[0] * gout[i];
    guuv[i] = 2.f * gout[i];
  }
  grot[0] += 2.f * (gout[0] * uv[0] + gout[1] * uv[1] + gout[2] * uv[2]);
  // uuv = q x uv : gq += uv x guuv ; guv += guuv x q
  cross3f(uv, guuv, t);
  for (int i = 0; i < 3; ++i) grot[1 + i] += t[i];
  cross3f(guuv, q, t);
  for (int i = 0; i < 3; ++i) guv[i] += t[i];
  // uv = q x v : gq += v x guv ; gv += guv x q
  cross3f(v, guv, t);
  for (int i = 0; i < 3; ++i) grot[1 + i] += t[i];
  cross3f(guv, q, t);
  for (int i = 0; i < 3; ++i) gv[i] += t[i];
}

// iq = conj(rot) / (|rot| + 1e-12)
UD_DEV void inv_quat_bwd(const float rot[4], const float giq[4], float grot[4]) {
  float nr = sqrtf(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2] + rot[3] * rot[3]);
  float n = nr + 1e-12f;
  const float sg[4] = {1.f, -1.f, -1.f, -1.f};
  float gn = 0.f;
  for (int i = 0; i < 4; ++i) {
    grot[i] += sg[i] * giq[i] / n;
    gn -= giq[i] * (sg[i] * rot[i]) / (n * n);
  }
  for (int i = 0; i < 4; ++i) grot[i] += gn * rot[i] / nr;
}

UD_DEV float sgnf(float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); }

UD_DEV void sdf_box_bwd(const float size[3], const float p[3], float g, float gsize[3], float gp[3]) {
  float q[3], raw[3];
  for (int i = 0; i < 3; ++i) {
    raw[i] = fabsf(p[i]) - size[i];
    q[i] = raw[i] >= 0.f ? raw[i] : 0.f;
  }
  float len = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + 1e-12f);
  float gq[3] = {g * q[0] / len, g * q[1] / len, g * q[2] / len};
  // "+ min(max(q), 0)": passes a gradient only when the selected q is exactly 0
  int sel = q[1] > q[2] ? 1 : 2;
  if (q[0] > q[sel]) sel = 0;
  if (q[sel] <= 0.f) gq[sel] += g;
  for (int i = 0; i < 3; ++i)
    if (raw[i] >= 0.f) {
      gp[i] += gq[i] * sgnf(p[i]);
      gsize[i] -= gq[i];
    }
}

UD_DEV void sdf_container_bwd(const float size[3], const float p[3], float g, float gsize[3], float gp[3]) {
  const float r = size[0], h = size[1];
  float w = sqrtf(r * r - h * h);
  float q0 = sqrtf(p[0] * p[0] + p[2] * p[2] + 1e-12f), q1 = p[1];
  float gq0 = 0.f, gq1 = 0.f, gw = 0.f, gr = 0.f, gh = 0.f, gt = -g;
  if (h * q0 < w * q1) {
    float a = q0 - w, b = q1 - h;
    float L = sqrtf(a * a + b * b + 1e-12f);
    gq0 = g * a / L;
    gw = -gq0;
    gq1 = g * b / L;
    gh = -gq1;
  } else {
    float L = sqrtf(q0 * q0 + q1 * q1 + 1e-12f);
    float sg = sgnf(L - r);
    gr = -g * sg;
    gq0 = g * sg * q0 / L;
    gq1 = g * sg * q1 / L;
  }
  gr += gw * r / w;
  gh -= gw * h / w;
  gsize[0] += gr;
  gsize[1] += gh;
  gsize[2] += gt;
  gp[0] += gq0 * p[0] / q0;
  gp[2] += gq0 * p[2] / q0;
  gp[1] += gq1;
}

UD_DEV void sdf_local_bwd(int kind, const float size[3], const float p[3], float g, float gsize[3], float gp[3]) {
  if (g == 0.f) return;
  if (kind == 0) sdf_box_bwd(size, p, g, gsize, gp);
  else sdf_container_bwd(size, p, g, gsize, gp);
}

// Reverse of collide_cell: given v_in (the velocity BEFORE this primitive) and gout = cotangent of the
// velocity after it, returns gv_in (overwritten) and accumulates the primitive's cotangents.
UD_DEV void collide_cell_bwd(int kind, float dt, const float gpos[3], const PrimIn<float>& pr, const float vin[3],
                             const float gout[3], float gvin[3], PrimGrad& pg) {
  // ---- recompute the forward
  float iq[4], rel[3], gp[3];
  inv_quat(pr.rot_f, iq);
  for (int i = 0; i < 3; ++i) rel[i] = gpos[i] - pr.pos_f[i];
  qrot(iq, rel, gp);
  float dist = sdf_local(kind, pr.size, gp);
  float ex = expf((-dist) * pr.softness);
  float infl = ex <= 1.f ? ex : 1.f;
  const float d = 1.e-6f, cfd = 0.5f / d;
  float nl[3];
  for (int a = 0; a < 3; ++a) {
    float inc[3] = {gp[0], gp[1], gp[2]}, dec[3] = {gp[0], gp[1], gp[2]};
    inc[a] = inc[a] + d;
    dec[a] = dec[a] + (-d);
    nl[a] = cfd * (sdf_local(kind, pr.size, inc) - sdf_local(kind, pr.size, dec));
  }
  float nlen = sqrtf(nl[0] * nl[0] + nl[1] * nl[1] + nl[2] * nl[2] + 1e-12f);
  float nn[3] = {nl[0] / nlen, nl[1] / nlen, nl[2] / nlen};
  float D[3], np[3], cv[3], iv[3];
  qrot(pr.rot_f, nn, D);
  qrot(pr.rot_f1, gp, np);
  for (int i = 0; i < 3; ++i) {
    cv[i] = ((np[i] + pr.pos_f1[i]) - gpos[i]) / dt;
    iv[i] = vin[i] - cv[i];
  }
  float nc = iv[0] * D[0] + iv[1] * D[1] + iv[2] * D[2];
  float ncm = nc <= 0.f ? nc : 0.f;
  float vt[3] = {iv[0] - ncm * D[0], iv[1] - ncm * D[1], iv[2] - ncm * D[2]};
  float vt2 = vt[0] * vt[0] + vt[1] * vt[1] + vt[2] * vt[2];
  float vtn = sqrtf(vt2 + 1e-12f);
  float fr_raw = vtn + nc * pr.friction;
  float fr = fr_raw >= 1e-12f ? fr_raw : 1e-12f;
  bool flag = nc < 0.f && sqrtf(vt2) > 1e-12f;
  // ---- reverse:  out = cv + iv (1 - infl) + vsel infl
  // giv = gout (1 - infl) + ex,  gcv = gout - giv = gout infl - ex   (ex collects the vsel path; written
  // this way the small-influence case does not cancel 1 - (1 - infl) in fp32)
  float gcv[3], ex_[3], gvt[3], gD[3] = {0.f, 0.f, 0.f};
  float ginfl = 0.f, gnc = 0.f, gfric = 0.f;
  for (int i = 0; i < 3; ++i) {
    float vsel = flag ? vt[i] / vtn * fr : vt[i];
    ginfl += gout[i] * (vsel - iv[i]);
  }
  if (flag) {
    // vsel_i = vt_i * s, s = fr / vtn
    float sc = fr / vtn, gs = 0.f;
    for (int i = 0; i < 3; ++i) {
      float gvs = gout[i] * infl;
      gvt[i] = gvs * sc;
      gs += gvs * vt[i];
    }
    float gfr = gs / vtn;
    float gvtn = -gs * fr / (vtn * vtn);
    if (fr_raw >= 1e-12f) {
      gvtn += gfr;
      gnc += gfr * pr.friction;
      gfric += gfr * nc;
    }
    for (int i = 0; i < 3; ++i) gvt[i] += gvtn * vt[i] / vtn;
  } else {
    for (int i = 0; i < 3; ++i) gvt[i] = gout[i] * infl;
  }
  // vt = iv - ncm D ; ncm = min(nc, 0) ; nc = iv . D
  float gncm = 0.f;
  for (int i = 0; i < 3; ++i) {
    ex_[i] = gvt[i];
    gncm -= gvt[i] * D[i];
    gD[i] -= ncm * gvt[i];
  }
  if (nc <= 0.f) gnc += gncm;
  for (int i = 0; i < 3; ++i) {
    ex_[i] += gnc * D[i];
    gD[i] += gnc * iv[i];
  }
  // iv = v - cv
  for (int i = 0; i < 3; ++i) {
    gvin[i] = gout[i] * (1.f - infl) + ex_[i];
    gcv[i] = gout[i] * infl - ex_[i];
  }
  // cv = (np + pos_f1 - gpos) / dt ; np = qrot(rot_f1, gp)
  float gnp[3], ggp[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < 3; ++i) {
    gnp[i] = gcv[i] / dt;
    pg.g[7 + i] += gnp[i];
  }
  qrot_bwd(pr.rot_f1, gp, gnp, &pg.g[10], ggp);
  // D = qrot(rot_f, nn) ; nn = nl / nlen
  float gnn[3] = {0.f, 0.f, 0.f};
  qrot_bwd(pr.rot_f, nn, gD, &pg.g[3], gnn);
  float dotn = gnn[0] * nl[0] + gnn[1] * nl[1] + gnn[2] * nl[2];
  float gsize[3] = {0.f, 0.f, 0.f};
  for (int a = 0; a < 3; ++a) {
    float gnl = gnn[a] / nlen - dotn * nl[a] / (nlen * nlen * nlen);
    if (gnl != 0.f) {
      float inc[3] = {gp[0], gp[1], gp[2]}, dec[3] = {gp[0], gp[1], gp[2]};
      inc[a] = inc[a] + d;
      dec[a] = dec[a] + (-d);
      sdf_local_bwd(kind, pr.size, inc, cfd * gnl, gsize, ggp);
      sdf_local_bwd(kind, pr.size, dec, -cfd * gnl, gsize, ggp);
    }
  }
  // infl = min(exp(-dist softness), 1) ; dist = sdf(size, gp)
  if (ex <= 1.f) sdf_local_bwd(kind, pr.size, gp, -pr.softness * infl * ginfl, gsize, ggp);
  for (int i = 0; i < 3; ++i) pg.g[14 + i] += gsize[i];
  pg.g[17] += gfric;
  // gp = qrot(iq, rel) ; rel = gpos - pos_f ; iq = inv_quat(rot_f)
  float giq[4] = {0.f, 0.f, 0.f, 0.f}, grel[3] = {0.f, 0.f, 0.f};
  qrot_bwd(iq, rel, ggp, giq, grel);
  for (int i = 0; i < 3; ++i) pg.g[i] -= grel[i];
  inv_quat_bwd(pr.rot_f, giq, &pg.g[3]);
}

// Reverse of position_control_cell (the mask carries no gradient).
UD_DEV void position_control_cell_bwd(int kind, float dt, const float gpos[3], const PrimIn<float>& pr,
                                      const float gout[3], float gvin[3], PrimGrad& pg) {
  float iq[4], rel[3], gp[3];
  inv_quat(pr.rot_f, iq);
  for (int i = 0; i < 3; ++i) rel[i] = gpos[i] - pr.pos_f[i];
  qrot(iq, rel, gp);
  float dist = sdf_local(kind, pr.size, gp);
  bool mask = dist < pr.size[0] * 1.5f;
  for (int i = 0; i < 3; ++i) {
    gvin[i] = mask ? 0.f : gout[i];
    if (mask) pg.g[18 + i] += gout[i] / dt;
  }
}

}  // namespace ud
