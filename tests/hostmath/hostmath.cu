// TEST-ONLY host build of the __host__ __device__ math headers of unidom_b200/csrc, so the per-particle
// and per-cell arithmetic (SVD, constitutive model and its reverse, collider and its reverse) can be
// checked against the CPU oracle without a GPU.  Never part of the product library.
#include "../../unidom_b200/csrc/mpm_cell.cuh"

using namespace ud;

static MpmConst make_k(const double* c) {
  // c = {dt, dx, inv_dx, p_mass, p_vol}
  MpmConst k{};
  k.dt = (float)c[0];
  k.dx = (float)c[1];
  k.inv_dx = (float)c[2];
  k.p_mass = (float)c[3];
  k.c_stress_mul = (float)(-c[0] * c[4] * 4);
  k.c_stress_div = (float)(c[1] * c[1]);
  k.sig_lo = (float)(1 - 2.5e-2 * 10);
  k.sig_hi = (float)(1 + 4.5e-3 * 100);
  return k;
}

static void fill_prim(const float* p, float softness, PrimIn<float>& pr) {
  for (int j = 0; j < 3; ++j) pr.pos_f[j] = p[j];
  for (int j = 0; j < 4; ++j) pr.rot_f[j] = p[3 + j];
  for (int j = 0; j < 3; ++j) pr.pos_f1[j] = p[7 + j];
  for (int j = 0; j < 4; ++j) pr.rot_f1[j] = p[10 + j];
  for (int j = 0; j < 3; ++j) pr.size[j] = p[14 + j];
  pr.friction = p[17];
  for (int j = 0; j < 3; ++j) pr.v_f[j] = p[18 + j];
  pr.softness = softness;
}

extern "C" {

void hm_svd3(int n, const float* A, float* U, float* s, float* Vt) {
  for (int i = 0; i < n; ++i) {
    Mat3 a, u, vt;
    for (int c = 0; c < 9; ++c) a.m[c] = A[9 * i + c];
    svd3(a, u, s + 3 * i, vt);
    for (int c = 0; c < 9; ++c) {
      U[9 * i + c] = u.m[c];
      Vt[9 * i + c] = vt.m[c];
    }
  }
}

// warm-started SVD (the previous substep's V^T as the starting basis), as k_p2g chains it over the substeps
void hm_svd3_warm(int n, const float* A, const float* Vt0, float* U, float* s, float* Vt) {
  for (int i = 0; i < n; ++i) {
    Mat3 a, u, vt;
    float v0[9];
    for (int c = 0; c < 9; ++c) {
      a.m[c] = A[9 * i + c];
      v0[c] = Vt0[9 * i + c];
    }
    svd3_ws(a, u, s + 3 * i, vt, true, v0);
    for (int c = 0; c < 9; ++c) {
      U[9 * i + c] = u.m[c];
      Vt[9 * i + c] = vt.m[c];
    }
  }
}

void hm_svd3_hestenes(int n, const float* A, float* U, float* s, float* Vt) {
  const float none[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  for (int i = 0; i < n; ++i) {
    Mat3 a, u, vt;
    for (int c = 0; c < 9; ++c) a.m[c] = A[9 * i + c];
    svd3_ws_hestenes(a, u, s + 3 * i, vt, false, none);
    for (int c = 0; c < 9; ++c) {
      U[9 * i + c] = u.m[c];
      Vt[9 * i + c] = vt.m[c];
    }
  }
}

void hm_constitutive(int n, const double* consts, const float* C, const float* F, float mu_s, float la_s,
                     const float* h, const int* material, const float* gA, const float* gF2, float* F2,
                     float* affine, float* gC, float* gF, float* gmu, float* gla) {
  MpmConst k = make_k(consts);
  for (int i = 0; i < n; ++i) {
    Mat3 c, f, ga, gf2, gc, gf;
    for (int j = 0; j < 9; ++j) {
      c.m[j] = C[9 * i + j];
      f.m[j] = F[9 * i + j];
      ga.m[j] = gA[9 * i + j];
      gf2.m[j] = gF2[9 * i + j];
    }
    Consti o;
    constitutive_fwd(k, c, f, mu_s, la_s, h[i], material[i], o);
    constitutive_bwd(k, c, f, o, ga, gf2, gc, gf, gmu[i], gla[i]);
    for (int j = 0; j < 9; ++j) {
      F2[9 * i + j] = o.F2.m[j];
      affine[9 * i + j] = o.affine.m[j];
      gC[9 * i + j] = gc.m[j];
      gF[9 * i + j] = gf.m[j];
    }
  }
}

// The adjoint's plastic fast path (plastic_affine + constitutive_bwd_plastic: same formulas in the SVD frame)
void hm_constitutive_plastic(int n, const double* consts, const float* C, const float* F, float mu_s, float la_s,
                             const float* h, const float* gA, const float* gF2, float* affine, float* gC, float* gF,
                             float* gmu, float* gla) {
  MpmConst k = make_k(consts);
  for (int i = 0; i < n; ++i) {
    Mat3 c, f, ga, gf2, gc, gf, aff;
    for (int j = 0; j < 9; ++j) {
      c.m[j] = C[9 * i + j];
      f.m[j] = F[9 * i + j];
      ga.m[j] = gA[9 * i + j];
      gf2.m[j] = gF2[9 * i + j];
    }
    Consti o;
    constitutive_pre(k, c, f, mu_s, la_s, h[i], 2, o);
    svd3(o.F1, o.U, o.s, o.Vt);
    plastic_affine(k, c, o.U, o.s, mu_s, la_s, h[i], aff);
    constitutive_bwd_plastic(k, c, f, o.U, o.s, o.Vt, mu_s, la_s, h[i], ga, gf2, gc, gf, gmu[i], gla[i]);
    for (int j = 0; j < 9; ++j) {
      affine[9 * i + j] = aff.m[j];
      gC[9 * i + j] = gc.m[j];
      gF[9 * i + j] = gf.m[j];
    }
  }
}

// The liquid fast path of P2G / P2G^T (constitutive_post_liquid + constitutive_bwd_liquid: no SVD, J = |det F1|)
void hm_constitutive_liquid(int n, const double* consts, const float* C, const float* F, const float* gA, const float* gF2,
                            float* F2, float* affine, float* gC, float* gF) {
  MpmConst k = make_k(consts);
  for (int i = 0; i < n; ++i) {
    Mat3 c, f, ga, gf2, gc, gf;
    for (int j = 0; j < 9; ++j) {
      c.m[j] = C[9 * i + j];
      f.m[j] = F[9 * i + j];
      ga.m[j] = gA[9 * i + j];
      gf2.m[j] = gF2[9 * i + j];
    }
    Consti o;
    constitutive_pre(k, c, f, 0.83f, 0.55f, 1.f, 0, o);
    constitutive_post_liquid(k, c, o);
    constitutive_bwd_liquid(k, c, f, ga, gf2, gc, gf);
    for (int j = 0; j < 9; ++j) {
      F2[9 * i + j] = o.F2.m[j];
      affine[9 * i + j] = o.affine.m[j];
      gC[9 * i + j] = gc.m[j];
      gF[9 * i + j] = gf.m[j];
    }
  }
}

// forward collide / position control on one cell
void hm_prim_fwd(int pos_control, int kind, float dt, const float* gpos, const float* prim, float softness,
                 const float* vin, float* vout) {
  PrimIn<float> pr;
  fill_prim(prim, softness, pr);
  float v[3] = {vin[0], vin[1], vin[2]};
  if (pos_control) position_control_cell(kind, dt, gpos, pr, v);
  else collide_cell(kind, dt, gpos, pr, v);
  for (int i = 0; i < 3; ++i) vout[i] = v[i];
}

// J^T gout by forward-mode duals (reference for the hand-written reverse): 24 inputs = vin(3) + prim(21)
void hm_prim_vjp_dual(int pos_control, int kind, float dt, const float* gpos, const float* prim, float softness,
                      const float* vin, const float* gout, float* gvin, float* gprim) {
  for (int ch = 0; ch < 3; ++ch) {
    typedef Dual<8> D;
    auto sd = [&](float val, int idx) {
      D r;
      r.v = val;
      for (int i = 0; i < 8; ++i) r.d[i] = (idx == ch * 8 + i) ? 1.f : 0.f;
      return r;
    };
    PrimIn<D> pr;
    for (int j = 0; j < 3; ++j) pr.pos_f[j] = sd(prim[j], 3 + j);
    for (int j = 0; j < 4; ++j) pr.rot_f[j] = sd(prim[3 + j], 6 + j);
    for (int j = 0; j < 3; ++j) pr.pos_f1[j] = sd(prim[7 + j], 10 + j);
    for (int j = 0; j < 4; ++j) pr.rot_f1[j] = sd(prim[10 + j], 13 + j);
    for (int j = 0; j < 3; ++j) pr.size[j] = sd(prim[14 + j], 17 + j);
    pr.friction = sd(prim[17], 20);
    for (int j = 0; j < 3; ++j) pr.v_f[j] = sd(prim[18 + j], 21 + j);
    pr.softness = softness;
    D v[3] = {sd(vin[0], 0), sd(vin[1], 1), sd(vin[2], 2)};
    if (pos_control) position_control_cell(kind, dt, gpos, pr, v);
    else collide_cell(kind, dt, gpos, pr, v);
    for (int i = 0; i < 8; ++i) {
      int id = ch * 8 + i;
      float g = gout[0] * v[0].d[i] + gout[1] * v[1].d[i] + gout[2] * v[2].d[i];
      if (id < 3) gvin[id] = g;
      else if (id < 24) gprim[id - 3] = g;
    }
  }
}

void hm_prim_vjp_hand(int pos_control, int kind, float dt, const float* gpos, const float* prim, float softness,
                      const float* vin, const float* gout, float* gvin, float* gprim) {
  PrimIn<float> pr;
  fill_prim(prim, softness, pr);
  PrimGrad pg;
  for (int i = 0; i < PRIM_NIN; ++i) pg.g[i] = 0.f;
  if (pos_control) position_control_cell_bwd(kind, dt, gpos, pr, gout, gvin, pg);
  else collide_cell_bwd(kind, dt, gpos, pr, vin, gout, gvin, pg);
  for (int i = 0; i < PRIM_NIN; ++i) gprim[i] = pg.g[i];
}

}  // extern "C"
