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

// mpm_simulator.py:283-313 for one cell with m > 0.  p = scattered momentum, m = scattered mass.
// `prim_of(q, pr)` fills primitive q and returns false when the primitive is to be skipped (the
// adjoint skips primitives whose influence on this cell is below 1e-12; the forward never skips).
template <class T, class PF>
UD_DEV void cell_update(const MpmConst& k, int ci, int cj, int ck, const T p[3], const T& m, const T& sfric,
                        PF prim_of, T v[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = p[i] / m + k.gdt[i];
  const float gpos[3] = {(float)ci * k.dx, (float)cj * k.dx, (float)ck * k.dx};
  for (int q = 0; q < k.n_prim; ++q) {
    PrimIn<T> pr;
    if (!prim_of(q, pr)) continue;
    if (k.pos_control)
      position_control_cell(k.sdf_kind, k.dt, gpos, pr, v);
    else
      collide_cell(k.sdf_kind, k.dt, gpos, pr, v);
  }
  // ground friction (:297-307)
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
  // boundary (:310-313); upper wall tests n_grid, not res
  const int cidx[3] = {ci, cj, ck};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    bool cond = (cidx[d] < 3 && s_val(v[d]) < 0.f) || (cidx[d] > k.n_grid - 3 && s_val(v[d]) > 0.f);
    if (cond) v[d] = s_const(v[d], 0.f);
  }
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

}  // namespace ud
