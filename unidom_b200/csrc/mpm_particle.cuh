// Per-particle MLS-MPM math: stencil, constitutive model forward and hand-derived reverse.
// Reference: DaXBench/daxbench/core/engine/mpm_simulator.py:227-268 (forward); the reverse is what
// jax.grad of the recomputed substep produces (mpm_simulator.py:339-362) with the SVD VJP of
// svd_safe_batch.py:65-102.
#pragma once
#include "common.cuh"

namespace ud {

struct MpmConst {
  int B, n, S, N;            // envs, particles/env, substeps, B*n
  int n_pad, N_pad;          // n rounded up to a whole warp tile (32), B*n_pad: extent of the sorted tile arrays
  int G, rx, ry, rz, n_grid; // cells/env, res, conf.n_grid
  int nbx, nby, nbz, NK;     // 4x4x4 blocks per axis, keys per env (= nbx*nby*nbz*64)
  float dt, dx, inv_dx, p_mass;
  float c_stress_mul;        // float(-dt*p_vol*4)           (mpm_simulator.py:267)
  float c_stress_div;        // float(dx**2)
  float gdt[3];              // float(dt)*float(gravity)     (:285)
  float sig_lo, sig_hi;      // float(1-2.5e-2*10), float(1+4.5e-3*100)  (:250)
  int n_prim, sdf_kind, pos_control, p2g_mode;
  int liquid_fast;           // UD_P2G_LIQUID_FAST: kernels with the SVD-free liquid path (template LIQ)
  int mark;  // development switch: 0 no block marks (timing only, wrong results), 1 every (segment,node), 2 corner nodes
};

// mul / sub without FMA contraction (the binning key must match the reference bit for bit)
UD_DEV float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
UD_DEV float sub_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}

struct Stencil {
  int base[3];
  float fx[3];
  float w[3][3];   // w[a][axis]: weight of offset a along axis
  float dw[3][3];  // d w[a][axis] / d fx[axis]
};

// mpm_simulator.py:233-235
UD_DEV void make_stencil(const float x[3], float inv_dx, Stencil& st) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float xs = mul_rn(x[d], inv_dx);
    int b = (int)sub_rn(xs, 0.5f);  // astype(int32): truncation toward zero
    float fx = sub_rn(xs, (float)b);
    st.base[d] = b;
    st.fx[d] = fx;
    float t0 = 1.5f - fx, t1 = fx - 1.f, t2 = fx - 0.5f;
    st.w[0][d] = 0.5f * (t0 * t0);
    st.w[1][d] = 0.75f - t1 * t1;
    st.w[2][d] = 0.5f * (t2 * t2);
    st.dw[0][d] = -t0;
    st.dw[1][d] = -2.f * t1;
    st.dw[2][d] = t2;
  }
}

// JAX index rules for one axis (negative wraps once, then out of range):
//   scatter: dropped  -> returns -1 ; gather: clamped
UD_DEV int idx_scatter(int i, int r) {
  if (i < 0) i += r;
  return (i < 0 || i >= r) ? -1 : i;
}
UD_DEV int idx_gather(int i, int r) {
  if (i < 0) i += r;
  return i < 0 ? 0 : (i >= r ? r - 1 : i);
}

struct Consti {
  Mat3 F1, U, Vt, F2, D, affine;  // D = F2 - U Vt
  float s[3], sc[3], J, mu, la, hc;
  bool plastic, liquid;
};

// mpm_simulator.py:238-245: F1 = (I + dt C) F and the per-particle Lame parameters
UD_DEV void constitutive_pre(const MpmConst& k, const Mat3& C, const Mat3& F, float mu_s, float la_s, float h,
                             int material, Consti& o) {
  o.liquid = material == 0;
  o.plastic = material == 2;
  Mat3 A;
#pragma unroll
  for (int i = 0; i < 9; ++i) A.m[i] = k.dt * C.m[i];
  A.m[0] += 1.f;
  A.m[4] += 1.f;
  A.m[8] += 1.f;
  o.F1 = mat_mul(A, F);
  o.hc = fminf(fmaxf(h, 0.1f), 5.f);
  o.mu = o.liquid ? 0.f : mu_s * o.hc;
  o.la = o.liquid ? 1.f : la_s * o.hc;
}

// mpm_simulator.py:249-268 given the SVD (o.U, o.s, o.Vt) of o.F1: plastic clip, J, F2, stress, affine
// all_plastic (warp-uniform in the kernels): every lane's particle is plastic, so the stress is formed in the frame of
// the SVD, M = (F2 - R) F2^T = U diag((sc - 1) sc) U^T (the formula plastic_affine and the adjoint already use): two
// 3x3 products fewer, and neither R nor D is formed (o.D is left unset).
UD_DEV void constitutive_post(const MpmConst& k, const Mat3& C, Consti& o, bool all_plastic = false) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o.sc[i] = o.plastic ? fminf(fmaxf(o.s[i], k.sig_lo), k.sig_hi) : o.s[i];
  o.J = o.sc[0] * o.sc[1] * o.sc[2];
  if (all_plastic) {
    Mat3 Us, Um;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Us(i, j) = o.U(i, j) * o.sc[j];
        Um(i, j) = Us(i, j) * (o.sc[j] - 1.f);
      }
    o.F2 = mat_mul(Us, o.Vt);
    const float iso = o.la * o.J * (o.J - 1.f);
    const float cs = k.c_stress_mul / k.c_stress_div;
    const float tm = 2.f * o.mu;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i; j < 3; ++j) {
        const float Mij = Um(i, 0) * o.U(j, 0) + Um(i, 1) * o.U(j, 1) + Um(i, 2) * o.U(j, 2);
        const float st = tm * Mij + (i == j ? iso : 0.f);
        o.affine(i, j) = cs * st + k.p_mass * C(i, j);
        if (i != j) o.affine(j, i) = cs * st + k.p_mass * C(j, i);
      }
    return;
  }
  Mat3 R = mat_mul(o.U, o.Vt);
  if (o.plastic) {
    Mat3 Us;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Us(i, j) = o.U(i, j) * o.sc[j];
    o.F2 = mat_mul(Us, o.Vt);
  } else {
    o.F2 = o.F1;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) o.D.m[i] = o.F2.m[i] - R.m[i];
  Mat3 M = mat_mul_nt(o.D, o.F2);
  float iso = o.la * o.J * (o.J - 1.f);
  const float cs = k.c_stress_mul / k.c_stress_div;  // (-dt*p_vol*4) / dx^2, one rounding earlier than :267
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float st = 2.f * o.mu * M(i, j) + (i == j ? iso : 0.f);
      o.affine(i, j) = cs * st + k.p_mass * C(i, j);
    }
}

// Liquid particle (material 0: mu = 0, la = 1, no clip, F2 = F1; mpm_simulator.py:241-242): the deviatoric term is
// 2 * 0 * M = 0 and the only thing left of the SVD is J = prod(sig) = |det F1|, so neither pass factorises F1.
// Cotangents: dU = dV = 0 and d sig_i = gJ prod_{j != i} sig_j, for which the reference's SVD VJP
// (svd_safe_batch.py:95-100) reduces to U diag(d sig) Vt = gJ sign(det F1) cof(F1).
UD_DEV void constitutive_post_liquid(const MpmConst& k, const Mat3& C, Consti& o) {
  o.J = fabsf(det3_pivoted(o.F1));
  o.F2 = o.F1;
  const float iso = o.la * o.J * (o.J - 1.f);
  const float cs = k.c_stress_mul / k.c_stress_div;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) o.affine(i, j) = cs * (i == j ? iso : 0.f) + k.p_mass * C(i, j);
}
// affine of a liquid particle from its input state (the adjoint's gather phase)
UD_DEV void liquid_affine(const MpmConst& k, const Mat3& C, const Mat3& F, Mat3& affine) {
  Consti o;
  constitutive_pre(k, C, F, 0.f, 0.f, 1.f, 0, o);
  constitutive_post_liquid(k, C, o);
  affine = o.affine;
}
// Reverse of the liquid constitutive step; d/d(state.mu), d/d(state.lamda) are zero (both are overwritten by constants).
UD_DEV void constitutive_bwd_liquid(const MpmConst& k, const Mat3& C, const Mat3& F, const Mat3& gA, const Mat3& gF2out,
                                    Mat3& gC, Mat3& gF) {
  Mat3 A;
#pragma unroll
  for (int i = 0; i < 9; ++i) A.m[i] = k.dt * C.m[i];
  A.m[0] += 1.f;
  A.m[4] += 1.f;
  A.m[8] += 1.f;
  const Mat3 F1 = mat_mul(A, F);
  const float det = det3_pivoted(F1);
  const float J = fabsf(det);
  const float cs = k.c_stress_mul / k.c_stress_div;
  const float trS = gA.m[0] * cs + gA.m[4] * cs + gA.m[8] * cs;
  const float gJ = (2.f * J - 1.f) * trS;   // la = 1
  const float gd = det < 0.f ? -gJ : gJ;
  const Mat3 cof = mat_cofactor(F1);
  Mat3 gF1;
#pragma unroll
  for (int i = 0; i < 9; ++i) gF1.m[i] = gF2out.m[i] + gd * cof.m[i];
  const Mat3 gCf = mat_mul_nt(gF1, F);
  gF = mat_mul_tn(A, gF1);
#pragma unroll
  for (int i = 0; i < 9; ++i) gC.m[i] = k.p_mass * gA.m[i] + k.dt * gCf.m[i];
}

// mpm_simulator.py:238-268
UD_DEV void constitutive_fwd(const MpmConst& k, const Mat3& C, const Mat3& F, float mu_s, float la_s,
                             float h, int material, Consti& o) {
  constitutive_pre(k, C, F, mu_s, la_s, h, material, o);
  svd3(o.F1, o.U, o.s, o.Vt);
  constitutive_post(k, C, o);
}

// Plastic material only (F2 = U diag(sc) Vt, R = U Vt): the stress needs neither F1 nor Vt,
//   M = (F2 - R) F2^T = U diag((sc - 1) sc) U^T.
// Used by the adjoint's gather phase, which only needs `affine` (1 symmetric product instead of 4 matrix products).
UD_DEV void plastic_affine(const MpmConst& k, const Mat3& C, const Mat3& U, const float s[3], float mu_s, float la_s,
                           float h, Mat3& affine) {
  const float hc = fminf(fmaxf(h, 0.1f), 5.f);
  const float mu = mu_s * hc, la = la_s * hc;
  float sc[3], m[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    sc[i] = fminf(fmaxf(s[i], k.sig_lo), k.sig_hi);
    m[i] = (sc[i] - 1.f) * sc[i];
  }
  const float J = sc[0] * sc[1] * sc[2];
  const float iso = la * J * (J - 1.f);
  const float cs = k.c_stress_mul / k.c_stress_div;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) {
      const float Mij = U(i, 0) * m[0] * U(j, 0) + U(i, 1) * m[1] * U(j, 1) + U(i, 2) * m[2] * U(j, 2);
      const float st = 2.f * mu * Mij + (i == j ? iso : 0.f);
      affine(i, j) = cs * st + k.p_mass * C(i, j);
      if (i != j) affine(j, i) = cs * st + k.p_mass * C(j, i);
    }
}

// Reverse of constitutive_fwd for a plastic particle, in the frame of the SVD.  The same formulas as
// constitutive_bwd (chain through F2 = U diag(sc) Vt, D = F2 - R, M = D F2^T and the reference's SVD VJP), with every
// product by U / Vt that cancels carried out symbolically: with Gs = U^T gS U, G = 2 mu Gs, H = U^T gF2out V,
//   U^T gD  V = G diag(sc)                         =: X
//   U^T gF2 V = G^T diag(sc - 1) + X + H           =: Y
//   U^T gU    = -X + Y diag(sc),   Vt gVt^T = -X^T + Y^T diag(sc),   gsc_i = gJ prod_{j != i} sc_j + Y_ii
//   <gS, M>   = sum_i (sc_i - 1) sc_i Gs_ii
// 8 matrix products instead of 17, and neither F1, F2, D nor R is formed.
UD_DEV void constitutive_bwd_plastic(const MpmConst& k, const Mat3& C, const Mat3& F, const Mat3& U, const float s[3],
                                     const Mat3& Vt, float mu_s, float la_s, float h, const Mat3& gA,
                                     const Mat3& gF2out, Mat3& gC, Mat3& gF, float& gmu_s, float& gla_s) {
  const float hc = fminf(fmaxf(h, 0.1f), 5.f);
  const float mu = mu_s * hc, la = la_s * hc;
  float sc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) sc[i] = fminf(fmaxf(s[i], k.sig_lo), k.sig_hi);
  const float J = sc[0] * sc[1] * sc[2];
  const float cs = k.c_stress_mul / k.c_stress_div;
  Mat3 gS;
#pragma unroll
  for (int i = 0; i < 9; ++i) gS.m[i] = gA.m[i] * cs;
  const float trS = gS.m[0] + gS.m[4] + gS.m[8];
  const Mat3 Gs = mat_mul(mat_mul_tn(U, gS), U);
  float gmu = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) gmu += (sc[i] - 1.f) * sc[i] * Gs(i, i);
  gmu_s = 2.f * gmu * hc;
  gla_s = J * (J - 1.f) * trS * hc;
  const float gJ = la * (2.f * J - 1.f) * trS;
  const Mat3 H = mat_mul_nt(mat_mul_tn(U, gF2out), Vt);
  const float tm = 2.f * mu;
  Mat3 UtdU, VdV;
  float gsc[3] = {gJ * sc[1] * sc[2], gJ * sc[0] * sc[2], gJ * sc[0] * sc[1]};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float Xij = tm * Gs(i, j) * sc[j], Xji = tm * Gs(j, i) * sc[i];
      const float Yij = tm * Gs(j, i) * (sc[j] - 1.f) + Xij + H(i, j);
      const float Yji = tm * Gs(i, j) * (sc[i] - 1.f) + Xji + H(j, i);
      UtdU(i, j) = -Xij + Yij * sc[j];
      VdV(i, j) = -Xji + Yji * sc[j];
      if (i == j) gsc[i] += Yij;
    }
  float gs[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    gs[i] = (s[i] > k.sig_lo && s[i] < k.sig_hi) ? gsc[i]
            : ((s[i] == k.sig_lo || s[i] == k.sig_hi) ? 0.5f * gsc[i] : 0.f);   // jnp.clip tie rule
  const Mat3 gF1 = svd3_bwd_rotated(U, s, Vt, UtdU, gs, VdV);
  const Mat3 gCf = mat_mul_nt(gF1, F);
  Mat3 A;
#pragma unroll
  for (int i = 0; i < 9; ++i) A.m[i] = k.dt * C.m[i];
  A.m[0] += 1.f;
  A.m[4] += 1.f;
  A.m[8] += 1.f;
  gF = mat_mul_tn(A, gF1);
#pragma unroll
  for (int i = 0; i < 9; ++i) gC.m[i] = k.p_mass * gA.m[i] + k.dt * gCf.m[i];
}

// Reverse of constitutive_fwd.  Inputs: cotangents of affine (gA) and of the output F (gF2out).
// Outputs: gC (adds p_mass*gA + dt*gF1 F^T), gF, and the per-particle contributions to the
// cotangents of state.mu / state.lamda.
UD_DEV void constitutive_bwd(const MpmConst& k, const Mat3& C, const Mat3& F, const Consti& o,
                             const Mat3& gA, const Mat3& gF2out, Mat3& gC, Mat3& gF, float& gmu_s,
                             float& gla_s) {
  const float cs = k.c_stress_mul / k.c_stress_div;
  Mat3 gS;  // cotangent of the unscaled stress
#pragma unroll
  for (int i = 0; i < 9; ++i) gS.m[i] = gA.m[i] * cs;
  Mat3 M = mat_mul_nt(o.D, o.F2);
  float trS = gS.m[0] + gS.m[4] + gS.m[8];
  float gmu = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) gmu += gS.m[i] * M.m[i];
  gmu *= 2.f;
  float gla = o.J * (o.J - 1.f) * trS;
  float gJ = o.la * (2.f * o.J - 1.f) * trS;
  gmu_s = o.liquid ? 0.f : gmu * o.hc;
  gla_s = o.liquid ? 0.f : gla * o.hc;
  // M = D F2^T : gD = gM F2 ; gF2 += gM^T D
  Mat3 gM;
#pragma unroll
  for (int i = 0; i < 9; ++i) gM.m[i] = 2.f * o.mu * gS.m[i];
  Mat3 gD = mat_mul(gM, o.F2);
  Mat3 gF2 = mat_mul_tn(gM, o.D);
#pragma unroll
  for (int i = 0; i < 9; ++i) gF2.m[i] += gD.m[i] + gF2out.m[i];
  // R = U Vt, D = F2 - R : gR = -gD
  Mat3 gU = mat_mul_nt(gD, o.Vt);   // gR Vt^T, negated below
  Mat3 gVt = mat_mul_tn(o.U, gD);   // U^T gR
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    gU.m[i] = -gU.m[i];
    gVt.m[i] = -gVt.m[i];
  }
  float gsc[3] = {gJ * o.sc[1] * o.sc[2], gJ * o.sc[0] * o.sc[2], gJ * o.sc[0] * o.sc[1]};
  Mat3 gF1 = mat_zero();
  if (o.plastic) {
    // F2 = U diag(sc) Vt
    Mat3 T = mat_mul_nt(gF2, o.Vt);  // gF2 Vt^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) gU(i, j) += T(i, j) * o.sc[j];
    Mat3 W = mat_mul_tn(o.U, gF2);  // U^T gF2
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gsc[i] += o.U(0, i) * T(0, i) + o.U(1, i) * T(1, i) + o.U(2, i) * T(2, i);
#pragma unroll
      for (int j = 0; j < 3; ++j) gVt(i, j) += o.sc[i] * W(i, j);
    }
  } else {
    gF1 = gF2;
  }
  float gs[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    gs[i] = o.plastic ? ((o.s[i] > k.sig_lo && o.s[i] < k.sig_hi) ? gsc[i]
                         : ((o.s[i] == k.sig_lo || o.s[i] == k.sig_hi) ? 0.5f * gsc[i] : 0.f))   // jnp.clip tie rule
                      : gsc[i];
  Mat3 dA = svd3_bwd(o.U, o.s, o.Vt, gU, gs, gVt);
#pragma unroll
  for (int i = 0; i < 9; ++i) gF1.m[i] += dA.m[i];
  // F1 = (I + dt C) F
  Mat3 gCf = mat_mul_nt(gF1, F);
  Mat3 A;
#pragma unroll
  for (int i = 0; i < 9; ++i) A.m[i] = k.dt * C.m[i];
  A.m[0] += 1.f;
  A.m[4] += 1.f;
  A.m[8] += 1.f;
  gF = mat_mul_tn(A, gF1);
#pragma unroll
  for (int i = 0; i < 9; ++i) gC.m[i] = k.p_mass * gA.m[i] + k.dt * gCf.m[i];
}

}  // namespace ud
