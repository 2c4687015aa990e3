// Shared device helpers for the unidom_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define UD_DEV __host__ __device__ __forceinline__

namespace ud {

// nan_to_num with JAX defaults (nan -> 0, +-inf -> +-FLT_MAX); mpm_simulator.py:377-381
UD_DEV float nan_to_num(float a) {
  if (a != a) return 0.f;
  if (isinf(a)) return a > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return a;
}

struct Mat3 {
  float m[9];  // row-major
  UD_DEV float& operator()(int r, int c) { return m[r * 3 + c]; }
  UD_DEV float operator()(int r, int c) const { return m[r * 3 + c]; }
};

UD_DEV Mat3 mat_mul(const Mat3& a, const Mat3& b) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
  return r;
}
// a * b^T
UD_DEV Mat3 mat_mul_nt(const Mat3& a, const Mat3& b) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r(i, j) = a(i, 0) * b(j, 0) + a(i, 1) * b(j, 1) + a(i, 2) * b(j, 2);
  return r;
}
// a^T * b
UD_DEV Mat3 mat_mul_tn(const Mat3& a, const Mat3& b) {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r(i, j) = a(0, i) * b(0, j) + a(1, i) * b(1, j) + a(2, i) * b(2, j);
  return r;
}
UD_DEV Mat3 mat_zero() {
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = 0.f;
  return r;
}

// Cofactor matrix cof(A)_ij = d det(A) / d A_ij (= det(A) A^-T).
UD_DEV Mat3 mat_cofactor(const Mat3& a) {
  Mat3 r;
  r(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
  r(0, 1) = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
  r(0, 2) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
  r(1, 0) = a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2);
  r(1, 1) = a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0);
  r(1, 2) = a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1);
  r(2, 0) = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1);
  r(2, 1) = a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2);
  r(2, 2) = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
  return r;
}
// det(A) by one elimination step with partial pivoting on column 0 and the 2x2 determinant of the rest: relative error
// ~ eps * cond(A) like an SVD's product of singular values (a plain cofactor expansion loses cond(A)^2).
UD_DEV float det3_pivoted(const Mat3& a) {
  float r0[3] = {a.m[0], a.m[1], a.m[2]}, r1[3] = {a.m[3], a.m[4], a.m[5]}, r2[3] = {a.m[6], a.m[7], a.m[8]};
  float sg = 1.f;
  if (fabsf(r1[0]) > fabsf(r0[0])) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { const float t = r0[i]; r0[i] = r1[i]; r1[i] = t; }
    sg = -sg;
  }
  if (fabsf(r2[0]) > fabsf(r0[0])) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { const float t = r0[i]; r0[i] = r2[i]; r2[i] = t; }
    sg = -sg;
  }
  const float inv = r0[0] != 0.f ? 1.f / r0[0] : 0.f;
  const float m1 = r1[0] * inv, m2 = r2[0] * inv;
  const float b11 = r1[1] - m1 * r0[1], b12 = r1[2] - m1 * r0[2], b21 = r2[1] - m2 * r0[1], b22 = r2[2] - m2 * r0[2];
  return sg * r0[0] * (b11 * b22 - b12 * b21);
}

// ---------------------------------------------------------------------------------------------
// 3x3 SVD  A = U diag(s) V^T,  s descending >= 0  (replaces jnp.linalg.svd, svd_safe_batch.py:51).
// One-sided (Hestenes) Jacobi: rotate column pairs of B = A V until orthogonal.  Fixed sweep count,
// branch-free rotations, all in registers.  Only the sign-invariant combinations U V^T, s and
// U diag(clip s) V^T are consumed downstream (mpm_simulator.py:253-265).
// Vt returned is V^T (the reference's "V"/"Vh").
// ---------------------------------------------------------------------------------------------
// fast (approximate, <= 2 ulp) division / sqrt / rsqrt for the Jacobi rotations: the iteration is
// self-correcting, only the final singular values and the U columns use IEEE sqrt / division.
// single-instruction (flush-to-zero) forms: the operands are O(1) quantities of a rotation, never denormal
UD_DEV float fast_div(float a, float b) {
#ifdef __CUDA_ARCH__
  float r;
  asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
#else
  return a / b;
#endif
}
UD_DEV float fast_sqrt(float a) {
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return sqrtf(a);
#endif
}
UD_DEV float fast_rsqrt(float a) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return 1.f / sqrtf(a);
#endif
}

// One Hestenes rotation of the column pair (p,q); returns the pair's normalised off-diagonal
// |b_p . b_q| / (|b_p| |b_q|) BEFORE the rotation (the orthogonality defect of U it removes).
//   t = 2g / (tau + sign(tau) sqrt(tau^2 + 4 g^2)),  tau = |b_q|^2 - |b_p|^2,  g = b_p . b_q
UD_DEV float jacobi_pair(float* bp, float* bq, float* vp, float* vq) {
  float alpha = bp[0] * bp[0] + bp[1] * bp[1] + bp[2] * bp[2];
  float beta = bq[0] * bq[0] + bq[1] * bq[1] + bq[2] * bq[2];
  float g2 = 2.f * (bp[0] * bq[0] + bp[1] * bq[1] + bp[2] * bq[2]);
  float tau = beta - alpha;
  float den = tau + copysignf(fast_sqrt(tau * tau + g2 * g2), tau);
  float t = den != 0.f ? fast_div(g2, den) : 0.f;
  float n2 = 1.f + t * t;
  float c = fast_rsqrt(n2);
  c = c * (1.5f - 0.5f * n2 * (c * c));  // one Newton step: V stays orthonormal to fp32 along warm-started chains
  float s = c * t;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float x = bp[i], y = bq[i];
    bp[i] = c * x - s * y;
    bq[i] = s * x + c * y;
    x = vp[i];
    y = vq[i];
    vp[i] = c * x - s * y;
    vq[i] = s * x + c * y;
  }
  return fabsf(g2) * fast_rsqrt(4.f * alpha * beta + 1e-37f);
}

UD_DEV void swap3(float* a, float* b) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float t = a[i];
    a[i] = b[i];
    b[i] = t;
  }
}

// Round-1 one-sided (Hestenes) variant, kept as the A/B partner and cross-check of svd3_ws below.
// Vt0 (optional, row-major V^T of a nearby matrix, e.g. the previous substep's): warm start.  The sweeps
// then start from B = A V0, whose columns are already nearly orthogonal, and typically stop after 1-2
// sweeps instead of 3-4.  Any orthogonal V0 is valid; the result is an SVD of A either way.
UD_DEV void svd3_ws_hestenes(const Mat3& A, Mat3& U, float s[3], Mat3& Vt, bool warm, const float (&Vt0)[9]) {
  // columns of B and V kept as separate arrays
  float b0[3] = {A(0, 0), A(1, 0), A(2, 0)}, b1[3] = {A(0, 1), A(1, 1), A(2, 1)},
        b2[3] = {A(0, 2), A(1, 2), A(2, 2)};
  float v0[3] = {1.f, 0.f, 0.f}, v1[3] = {0.f, 1.f, 0.f}, v2[3] = {0.f, 0.f, 1.f};
  if (warm) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      v0[i] = Vt0[i];
      v1[i] = Vt0[3 + i];
      v2[i] = Vt0[6 + i];
    }
    // re-orthonormalise V0 (Gram-Schmidt + cross product): every rotation is orthogonal only to ~1e-7, and
    // along a warm-started chain that drift would otherwise accumulate into U diag(s) Vt != A
    {
      float n0 = v0[0] * v0[0] + v0[1] * v0[1] + v0[2] * v0[2];
      float r0 = fast_rsqrt(n0);
      r0 = r0 * (1.5f - 0.5f * n0 * (r0 * r0));
      v0[0] *= r0; v0[1] *= r0; v0[2] *= r0;
      float d = v0[0] * v1[0] + v0[1] * v1[1] + v0[2] * v1[2];
      v1[0] -= d * v0[0]; v1[1] -= d * v0[1]; v1[2] -= d * v0[2];
      float n1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
      float r1 = fast_rsqrt(n1);
      r1 = r1 * (1.5f - 0.5f * n1 * (r1 * r1));
      v1[0] *= r1; v1[1] *= r1; v1[2] *= r1;
      float cx = v0[1] * v1[2] - v0[2] * v1[1], cy = v0[2] * v1[0] - v0[0] * v1[2], cz = v0[0] * v1[1] - v0[1] * v1[0];
      float sg = (cx * v2[0] + cy * v2[1] + cz * v2[2]) < 0.f ? -1.f : 1.f;
      v2[0] = sg * cx; v2[1] = sg * cy; v2[2] = sg * cz;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // B = A V0: column j of B = A v_j
      b0[i] = A(i, 0) * v0[0] + A(i, 1) * v0[1] + A(i, 2) * v0[2];
      b1[i] = A(i, 0) * v1[0] + A(i, 1) * v1[1] + A(i, 2) * v1[2];
      b2[i] = A(i, 0) * v2[0] + A(i, 1) * v2[1] + A(i, 2) * v2[2];
    }
  }
  // cyclic sweeps; quadratic convergence: once every normalised off-diagonal met in a sweep is
  // < 2e-4 the defect left after that sweep is O(1e-8), below fp32 resolution
#pragma unroll 1
  for (int sweep = 0; sweep < 6; ++sweep) {
    float t0 = jacobi_pair(b0, b1, v0, v1);
    float t1 = jacobi_pair(b0, b2, v0, v2);
    float t2 = jacobi_pair(b1, b2, v1, v2);
    if (fmaxf(t0, fmaxf(t1, t2)) < 2e-4f) break;
  }
  float n0 = b0[0] * b0[0] + b0[1] * b0[1] + b0[2] * b0[2];
  float n1 = b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2];
  float n2 = b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2];
  // sort descending (3-element network)
  if (n0 < n1) { swap3(b0, b1); swap3(v0, v1); float t = n0; n0 = n1; n1 = t; }
  if (n0 < n2) { swap3(b0, b2); swap3(v0, v2); float t = n0; n0 = n2; n2 = t; }
  if (n1 < n2) { swap3(b1, b2); swap3(v1, v2); float t = n1; n1 = n2; n2 = t; }
  s[0] = sqrtf(n0);
  s[1] = sqrtf(n1);
  s[2] = sqrtf(n2);
  float u0[3], u1[3], u2[3];
  float i0 = s[0] > 0.f ? 1.f / s[0] : 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) u0[i] = s[0] > 0.f ? b0[i] * i0 : (i == 0 ? 1.f : 0.f);
  if (s[1] > 1e-20f * s[0] && s[1] > 0.f) {
    float i1 = 1.f / s[1];
#pragma unroll
    for (int i = 0; i < 3; ++i) u1[i] = b1[i] * i1;
  } else {  // any unit vector orthogonal to u0
    float ax = fabsf(u0[0]), ay = fabsf(u0[1]), az = fabsf(u0[2]);
    float e[3] = {0.f, 0.f, 0.f};
    if (ax <= ay && ax <= az) e[0] = 1.f; else if (ay <= az) e[1] = 1.f; else e[2] = 1.f;
    float d = e[0] * u0[0] + e[1] * u0[1] + e[2] * u0[2];
    float w[3] = {e[0] - d * u0[0], e[1] - d * u0[1], e[2] - d * u0[2]};
    float inv = 1.f / sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) u1[i] = w[i] * inv;
  }
  if (s[2] > 1e-20f * s[0] && s[2] > 0.f) {
    float i2 = 1.f / s[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) u2[i] = b2[i] * i2;
  } else {
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
    u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
    u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    U(i, 0) = u0[i];
    U(i, 1) = u1[i];
    U(i, 2) = u2[i];
    Vt(0, i) = v0[i];
    Vt(1, i) = v1[i];
    Vt(2, i) = v2[i];
  }
}

#if defined(UD_SVD_STATS) && !defined(__CUDA_ARCH__)
static long g_svd_stats[64];   // host-only development counters (tests/hostmath with -DUD_SVD_STATS)
#endif
// One two-sided Jacobi rotation of the pair (p,q) of the symmetric Gram matrix S = B^T B (B = A V), r = third index:
// the same rotation angle as the Hestenes rotation of columns p, q of B, but the three dot products are the stored
// entries of S instead of being re-formed from B.  Rotates columns p, q of V along.  Returns the pair's normalised
// off-diagonal |S_pq| / sqrt(S_pp S_qq) BEFORE the rotation (the orthogonality defect of U it removes).
//   t = 2g / (tau + sign(tau) sqrt(tau^2 + 4 g^2)),  tau = S_qq - S_pp,  g = S_pq
UD_DEV float jacobi_sym_pair(float& spp, float& sqq, float& spq, float& spr, float& sqr, float* vp, float* vq) {
  const float g2 = 2.f * spq;
  const float tau = sqq - spp;
  const float den = tau + copysignf(fast_sqrt(tau * tau + g2 * g2), tau);
  const float t = den != 0.f ? fast_div(g2, den) : 0.f;
  const float n2 = 1.f + t * t;
  float c = fast_rsqrt(n2);
  c = c * (1.5f - 0.5f * n2 * (c * c));  // one Newton step: V stays orthonormal to fp32 along warm-started chains
  const float s = c * t;
  const float defect = fabsf(g2) * fast_rsqrt(4.f * spp * sqq + 1e-37f);
  const float tg = t * spq;
  spp -= tg;
  sqq += tg;
  spq = 0.f;
  const float x = spr, y = sqr;
  spr = c * x - s * y;
  sqr = s * x + c * y;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a = vp[i], b = vq[i];
    vp[i] = c * a - s * b;
    vq[i] = s * a + c * b;
  }
  return defect;
}

// Vt0 (optional, row-major V^T of a nearby matrix, e.g. the previous substep's): warm start.  The sweeps
// then start from B = A V0, whose columns are already nearly orthogonal, and typically stop after 1-2
// sweeps instead of 3-4.  Any orthogonal V0 is valid; the result is an SVD of A either way.
// Round 2: the sweeps run on the 3x3 Gram matrix S = (A V0)^T (A V0) (two-sided Jacobi, 6 + 9 live values) and B = A V
// is formed once at the end, instead of rotating the columns of B and re-forming three dot products per pair
// (one-sided Hestenes, kept below as svd3_ws_hestenes): ~35 instead of ~55 instructions per rotation.  The singular
// values of this path lie in [0.1, 10], so squaring the condition number costs nothing measurable in fp32
// (tests/test_hostmath_cpu.py compares both against LAPACK).
UD_DEV void svd3_ws(const Mat3& A, Mat3& U, float s[3], Mat3& Vt, bool warm, const float (&Vt0)[9]) {
#ifdef UD_SVD_HESTENES   // development A/B build (python -m unidom_b200.build --variant hestenes -DUD_SVD_HESTENES)
  svd3_ws_hestenes(A, U, s, Vt, warm, Vt0);
  return;
#endif
  float b0[3] = {A(0, 0), A(1, 0), A(2, 0)}, b1[3] = {A(0, 1), A(1, 1), A(2, 1)},
        b2[3] = {A(0, 2), A(1, 2), A(2, 2)};
  float v0[3] = {1.f, 0.f, 0.f}, v1[3] = {0.f, 1.f, 0.f}, v2[3] = {0.f, 0.f, 1.f};
  if (warm) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      v0[i] = Vt0[i];
      v1[i] = Vt0[3 + i];
      v2[i] = Vt0[6 + i];
    }
    // re-orthonormalise V0 (Gram-Schmidt + cross product): every rotation is orthogonal only to ~1e-7, and
    // along a warm-started chain that drift would otherwise accumulate into U diag(s) Vt != A
    {
      float n0 = v0[0] * v0[0] + v0[1] * v0[1] + v0[2] * v0[2];
      float r0 = fast_rsqrt(n0);
      r0 = r0 * (1.5f - 0.5f * n0 * (r0 * r0));
      v0[0] *= r0; v0[1] *= r0; v0[2] *= r0;
      float d = v0[0] * v1[0] + v0[1] * v1[1] + v0[2] * v1[2];
      v1[0] -= d * v0[0]; v1[1] -= d * v0[1]; v1[2] -= d * v0[2];
      float n1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
      float r1 = fast_rsqrt(n1);
      r1 = r1 * (1.5f - 0.5f * n1 * (r1 * r1));
      v1[0] *= r1; v1[1] *= r1; v1[2] *= r1;
      float cx = v0[1] * v1[2] - v0[2] * v1[1], cy = v0[2] * v1[0] - v0[0] * v1[2], cz = v0[0] * v1[1] - v0[1] * v1[0];
      float sg = (cx * v2[0] + cy * v2[1] + cz * v2[2]) < 0.f ? -1.f : 1.f;
      v2[0] = sg * cx; v2[1] = sg * cy; v2[2] = sg * cz;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // B = A V0: column j of B = A v_j
      b0[i] = A(i, 0) * v0[0] + A(i, 1) * v0[1] + A(i, 2) * v0[2];
      b1[i] = A(i, 0) * v1[0] + A(i, 1) * v1[1] + A(i, 2) * v1[2];
      b2[i] = A(i, 0) * v2[0] + A(i, 1) * v2[1] + A(i, 2) * v2[2];
    }
  }
  float s00 = b0[0] * b0[0] + b0[1] * b0[1] + b0[2] * b0[2];
  float s11 = b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2];
  float s22 = b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2];
  float s01 = b0[0] * b1[0] + b0[1] * b1[1] + b0[2] * b1[2];
  float s02 = b0[0] * b2[0] + b0[1] * b2[1] + b0[2] * b2[2];
  float s12 = b1[0] * b2[0] + b1[1] * b2[1] + b1[2] * b2[2];
  // cyclic sweeps; quadratic convergence: once every normalised off-diagonal met in a sweep is
  // < 2e-4 the defect left after that sweep is O(1e-8), below fp32 resolution
#pragma unroll 1
  for (int sweep = 0; sweep < 6; ++sweep) {
    const float t0 = jacobi_sym_pair(s00, s11, s01, s02, s12, v0, v1);
    const float t1 = jacobi_sym_pair(s00, s22, s02, s01, s12, v0, v2);
    const float t2 = jacobi_sym_pair(s11, s22, s12, s01, s02, v1, v2);
#if defined(UD_SVD_STATS) && !defined(__CUDA_ARCH__)
    ++g_svd_stats[1 + sweep];
    { float m = fmaxf(t0, fmaxf(t1, t2)); int b = m < 1e-6f ? 0 : m < 1e-5f ? 1 : m < 1e-4f ? 2 : m < 2e-4f ? 3 : m < 1e-3f ? 4 : m < 1e-2f ? 5 : 6; ++g_svd_stats[8 + sweep * 8 + b]; }
#endif
    if (fmaxf(t0, fmaxf(t1, t2)) < 2e-4f) break;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {  // B = A V
    b0[i] = A(i, 0) * v0[0] + A(i, 1) * v0[1] + A(i, 2) * v0[2];
    b1[i] = A(i, 0) * v1[0] + A(i, 1) * v1[1] + A(i, 2) * v1[2];
    b2[i] = A(i, 0) * v2[0] + A(i, 1) * v2[1] + A(i, 2) * v2[2];
  }
  float n0 = b0[0] * b0[0] + b0[1] * b0[1] + b0[2] * b0[2];
  float n1 = b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2];
  float n2 = b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2];
  {
    // The recursion on S carries ~1e-7 |S| of absolute rounding, which is a RELATIVE orthogonality defect of
    // 1e-7 / (s_i s_j) in U: harmless for the singular values this path sees, not for an ill-conditioned A.  Measure
    // the true defects of B = A V and, in the rare case they are visible in fp32, polish with one-sided sweeps.
    const float d01 = b0[0] * b1[0] + b0[1] * b1[1] + b0[2] * b1[2];
    const float d02 = b0[0] * b2[0] + b0[1] * b2[1] + b0[2] * b2[2];
    const float d12 = b1[0] * b2[0] + b1[1] * b2[1] + b1[2] * b2[2];
    const float thr2 = 9e-12f;  // (3e-6)^2
    if (d01 * d01 > thr2 * n0 * n1 || d02 * d02 > thr2 * n0 * n2 || d12 * d12 > thr2 * n1 * n2) {
#if defined(UD_SVD_STATS) && !defined(__CUDA_ARCH__)
      ++g_svd_stats[0];
#endif
#pragma unroll 1
      for (int sweep = 0; sweep < 4; ++sweep) {
        const float t0 = jacobi_pair(b0, b1, v0, v1);
        const float t1 = jacobi_pair(b0, b2, v0, v2);
        const float t2 = jacobi_pair(b1, b2, v1, v2);
        if (fmaxf(t0, fmaxf(t1, t2)) < 2e-4f) break;
      }
      n0 = b0[0] * b0[0] + b0[1] * b0[1] + b0[2] * b0[2];
      n1 = b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2];
      n2 = b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2];
    }
  }
  // sort descending (3-element network); along a warm-started chain the order is already right: one branch
  if (n0 < n1 || n1 < n2) {
    if (n0 < n1) { swap3(b0, b1); swap3(v0, v1); float t = n0; n0 = n1; n1 = t; }
    if (n0 < n2) { swap3(b0, b2); swap3(v0, v2); float t = n0; n0 = n2; n2 = t; }
    if (n1 < n2) { swap3(b1, b2); swap3(v1, v2); float t = n1; n1 = n2; n2 = t; }
  }
  // s_i = |b_i|, u_i = b_i / s_i through one refined reciprocal square root each (<= 2 ulp)
  float i0 = n0 > 0.f ? fast_rsqrt(n0) : 0.f, i1 = n1 > 0.f ? fast_rsqrt(n1) : 0.f, i2 = n2 > 0.f ? fast_rsqrt(n2) : 0.f;
  i0 = i0 * (1.5f - 0.5f * n0 * (i0 * i0));
  i1 = i1 * (1.5f - 0.5f * n1 * (i1 * i1));
  i2 = i2 * (1.5f - 0.5f * n2 * (i2 * i2));
  s[0] = n0 * i0;
  s[1] = n1 * i1;
  s[2] = n2 * i2;
  float u0[3], u1[3], u2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) u0[i] = s[0] > 0.f ? b0[i] * i0 : (i == 0 ? 1.f : 0.f);
  if (s[1] > 1e-20f * s[0] && s[1] > 0.f) {
#pragma unroll
    for (int i = 0; i < 3; ++i) u1[i] = b1[i] * i1;
  } else {  // any unit vector orthogonal to u0
    float ax = fabsf(u0[0]), ay = fabsf(u0[1]), az = fabsf(u0[2]);
    float e[3] = {0.f, 0.f, 0.f};
    if (ax <= ay && ax <= az) e[0] = 1.f; else if (ay <= az) e[1] = 1.f; else e[2] = 1.f;
    float d = e[0] * u0[0] + e[1] * u0[1] + e[2] * u0[2];
    float w[3] = {e[0] - d * u0[0], e[1] - d * u0[1], e[2] - d * u0[2]};
    float inv = 1.f / sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) u1[i] = w[i] * inv;
  }
  if (s[2] > 1e-20f * s[0] && s[2] > 0.f) {
#pragma unroll
    for (int i = 0; i < 3; ++i) u2[i] = b2[i] * i2;
  } else {
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
    u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
    u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    U(i, 0) = u0[i];
    U(i, 1) = u1[i];
    U(i, 2) = u2[i];
    Vt(0, i) = v0[i];
    Vt(1, i) = v1[i];
    Vt(2, i) = v2[i];
  }
}

UD_DEV void svd3(const Mat3& A, Mat3& U, float s[3], Mat3& Vt) {
  const float none[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  svd3_ws(A, U, s, Vt, false, none);
}

// Reference VJP of the SVD, svd_safe_batch.py:65-102, real square case:
//   dA = U diag(dS) Vt + U((J+J^T) o S_col) Vt + (U o S_col)(K+K^T) Vt
//   Fij = safe_inv(s_j^2 - s_i^2), zero diagonal; J = F o (U^T dU); K = F o (Vt dVt^T)
// (the L and projector terms vanish identically for real 3x3 input).
UD_DEV float safe_inv(float x) { return x / (x * x + 1e-12f); }

// Core of the VJP given U^T dU and Vt dVt^T.
UD_DEV Mat3 svd3_bwd_rotated(const Mat3& U, const float s[3], const Mat3& Vt, const Mat3& UtdU, const float dS[3],
                             const Mat3& VdV) {
  float s2[3] = {s[0] * s[0], s[1] * s[1], s[2] * s[2]};
  Mat3 M;  // diag(dS) + (J+J^T) o S_col + S_row-scaled (K+K^T)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (i == j) {
        M(i, j) = dS[i];
      } else {
        // safe_inv is odd and fp subtraction anti-commutes exactly: F_ji = -F_ij bit for bit, so each of the three
        // index pairs costs ONE IEEE division (the two calls as written are 6 distinct divisions after CSE)
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        const float f = safe_inv(s2[hi] - s2[lo]);   // F_{lo,hi}
        const float Fij = i < j ? f : -f, Fji = -Fij;
        float Jsym = Fij * UtdU(i, j) + Fji * UtdU(j, i);
        float Ksym = Fij * VdV(i, j) + Fji * VdV(j, i);
        // (J+J^T)*S broadcasts S over the last axis (column j); (U*S) scales column i of U,
        // i.e. row i of (K+K^T).
        M(i, j) = Jsym * s[j] + s[i] * Ksym;
      }
    }
  return mat_mul(mat_mul(U, M), Vt);
}

UD_DEV Mat3 svd3_bwd(const Mat3& U, const float s[3], const Mat3& Vt, const Mat3& dU, const float dS[3],
                     const Mat3& dVt) {
  return svd3_bwd_rotated(U, s, Vt, mat_mul_tn(U, dU), dS, mat_mul_nt(Vt, dVt));
}

}  // namespace ud
