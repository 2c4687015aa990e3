// Mass-spring cloth step (fwd + adjoint) for sm_100a.  One CTA per environment, one thread per cloth
// node: positions/velocities live in registers, neighbour positions are exchanged through shared
// memory, and all `substeps` (50) substeps of a sub-action run inside ONE launch -- the cloth state
// (512 nodes x 24 B) never round-trips HBM between substeps.
// Cloths with more than 1024 nodes (fold_tshirt: 3 573) run on a thread-block CLUSTER per environment
// (up to 8 CTAs x 1024 threads): every CTA keeps a full copy of the node positions that the owners
// update through distributed shared memory, the spring cotangents stay in the owner's shared memory
// and are read remotely, and the per-env norms are reduced across the cluster, barrier by barrier.
// Reference: DaXBench/daxbench/core/engine/cloth_simulator.py:163-180 (robot_step), :198-226 (grippers),
// :257-337 (step), :182-196 (norm_grad, re-normalises the cotangent 8x per substep in the adjoint).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/unidom_b200.h"
#include "mpm_internal.h"  // KScope instrumentation

namespace ud {

namespace cg = cooperative_groups;

struct ClothK {
  int B, P, S, threads, CL;   // threads per CTA, CTAs (cluster size) per env
  float dt, g, gdt, damp, max_v, small_num, mask_sum;
  int stiff_float;
};

static int g_cta_nodes = 1024;   // nodes per CTA before an env spills onto a cluster (A/B + test switch, <= 1024)
int cloth_tuning_cta_nodes(int v) {
  int o = g_cta_nodes;
  if (v >= 32 && v <= 1024) g_cta_nodes = v / 32 * 32;
  return o;
}

static bool cloth_fold(const ud_cloth_params* p, ClothK* k) {
  if (!p || p->num_envs < 1 || p->n_nodes < 1 || p->n_nodes > 8 * g_cta_nodes || p->substeps < 1) return false;
  if (!(p->dt > 0) || !(p->mask_sum > 0)) return false;
  k->B = p->num_envs;
  k->P = p->n_nodes;
  k->S = p->substeps;
  int per = g_cta_nodes;
  // beyond one CTA prefer 512-thread CTAs (128 registers per thread, no spills) while 8 of them cover the cloth
  if (p->n_nodes > per && per == 1024 && p->n_nodes <= 8 * 512) per = 512;
  k->CL = (p->n_nodes + per - 1) / per;
  k->threads = ((p->n_nodes + k->CL - 1) / k->CL + 31) / 32 * 32;
  k->dt = (float)p->dt;
  k->g = (float)p->gravity;
  k->gdt = (float)(p->gravity * p->dt);                 // jnp.array([0, gravity*dt, 0]) (:259)
  k->damp = expf((float)(-p->damping * p->dt));         // jnp.exp(-damping*dt) in f32 (:309)
  k->max_v = (float)p->max_v;
  k->small_num = (float)p->small_num;
  k->mask_sum = (float)p->mask_sum;
  k->stiff_float = p->stiffness_is_float;
  return true;
}

__device__ __forceinline__ float clipf(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }
// d clip(a, lo, hi) / d a with jnp.clip's tie rule: clip = minimum(hi, maximum(lo, a)) and lax.max/min split the
// cotangent evenly at a tie (a == lo or a == hi).  Nodes rest at y == 0 exactly after reset and the idle
// gripper sits at 1.0 exactly, so ties are hit in practice.
__device__ __forceinline__ float clip_grad(float a, float lo, float hi) {
  return (a > lo && a < hi) ? 1.f : ((a == lo || a == hi) ? 0.5f : 0.f);
}
// a / b given rb = RN(1 / b): one Newton correction on the remainder reproduces the correctly rounded IEEE quotient
// (normal range), so several divisions by the same denominator share one reciprocal.  The cloth is chaotic; keeping
// the reference's exact quotients keeps the trajectories on the reference's branch for as long as possible.
__device__ __forceinline__ float div_rn(float a, float b, float rb) {
  const float q = a * rb;
  return __fmaf_rn(__fmaf_rn(-q, b, a), rb, q);
}
__device__ __forceinline__ float nan0(float a);
// norm_grad's backward on one component: nan_to_num(g / n) / mask_sum (:190-192), n and mask_sum shared by many
// components.  rn = RN(1/n); when n == 0 the reciprocal is inf and the true division decides (0/0 -> NaN -> 0,
// x/0 -> inf -> FLT_MAX), exactly as the reference's arithmetic does.
__device__ __forceinline__ float norm_div(float g, float n, float rn, float ms, float rms) {
  const float q = isinf(rn) ? g / n : div_rn(g, n, rn);
  return div_rn(nan0(q), ms, rms);
}
__device__ __forceinline__ float nan0(float a) {   // jnp.nan_to_num: NaN -> 0, +-inf -> +-FLT_MAX
  const float c = fminf(fmaxf(a, -3.4028234663852886e38f), 3.4028234663852886e38f);
  return a == a ? c : 0.f;
}

// One env = one CTA (CLUSTER = false) or one thread-block cluster (CLUSTER = true).
template <bool CLUSTER>
struct Team {
  __device__ __forceinline__ static int rank() { return CLUSTER ? (int)cg::this_cluster().block_rank() : 0; }
  __device__ __forceinline__ static int size() { return CLUSTER ? (int)cg::this_cluster().num_blocks() : 1; }
  __device__ __forceinline__ static int env() { return CLUSTER ? (int)(blockIdx.x / cg::this_cluster().num_blocks()) : (int)blockIdx.x; }
  __device__ __forceinline__ static void sync() {
    if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
  }
  // publish this node's position into the position table of every CTA of the env
  __device__ __forceinline__ static void publish(float* xs, int n, const float x[3]) {
    if (CLUSTER) {
      cg::cluster_group cl = cg::this_cluster();
      for (unsigned r = 0; r < cl.num_blocks(); ++r) {
        float* d = cl.map_shared_rank(xs, r);
        d[3 * n] = x[0];
        d[3 * n + 1] = x[1];
        d[3 * n + 2] = x[2];
      }
    } else {
      xs[3 * n] = x[0];
      xs[3 * n + 1] = x[1];
      xs[3 * n + 2] = x[2];
    }
  }
};

// env-wide sums of two values (all threads of the env must call); result broadcast to every thread.
// red: [64] per CTA; redc: [2][16] per CTA, double-buffered by `par` (a fast CTA may already be writing the next
// reduction's partials while a slow one still reads this one's).  Fixed summation order: deterministic.
template <bool CLUSTER>
__device__ __forceinline__ void team_sum2(float& a, float& b, float* red, float* redc, int& par) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, off);
    b += __shfl_down_sync(0xffffffffu, b, off);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) {
    red[wid] = a;
    red[32 + wid] = b;
  }
  __syncthreads();
  // second stage: every warp folds the <= 32 warp partials with the same xor butterfly, so all threads hold the
  // bit-identical total (fixed tree: deterministic) after 2 loads + 10 shuffles instead of 2 * nw loads
  float sa = lane < nw ? red[lane] : 0.f, sb = lane < nw ? red[32 + lane] : 0.f;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, off);
    sb += __shfl_xor_sync(0xffffffffu, sb, off);
  }
  if (CLUSTER) {
    cg::cluster_group cl = cg::this_cluster();
    const int nb = (int)cl.num_blocks(), me = (int)cl.block_rank();
    if ((int)threadIdx.x < nb) {
      float* r = cl.map_shared_rank(redc, threadIdx.x);
      r[par * 16 + me] = sa;
      r[par * 16 + 8 + me] = sb;
    }
    cl.sync();
    sa = 0.f;
    sb = 0.f;
    for (int r = 0; r < nb; ++r) {
      sa += redc[par * 16 + r];
      sb += redc[par * 16 + 8 + r];
    }
    par ^= 1;
  }
  a = sa;
  b = sb;
}

struct Sub {  // forward intermediates of one substep for one node (kept for the reverse)
  float vg[3];        // v after the gravity kick
  float fr[3];        // spring force sum + gravity (before friction)
  float f[3];         // force after friction
  float muF, sV, sF;
  bool fmask, dyn, stat, zero, nonz;
  float v0[3];        // after damping, before grippers
  bool m0, m1;        // gripper masks
  float x1[3], v1[3]; // after gripper 0
  float x2[3], v2[3]; // after gripper 1
};

// one forward substep for one node; xs = shared positions of all nodes (already synchronised)
__device__ __forceinline__ void cloth_substep(const ClothK& k, const float* __restrict__ xs, const int nbr[8],
                                              const float L0[8], const float iL0[8], float stiff, float mu,
                                              const float a0[4],
                                              const float a1[4], const float ps0[4], const float ps1[4],
                                              float x[3], float v[3], Sub& o) {
  o.vg[0] = v[0];
  o.vg[1] = v[1] - k.gdt;
  o.vg[2] = v[2];
  float fs[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (nbr[q] < 0) continue;
    float r0 = xs[3 * nbr[q]] - x[0], r1 = xs[3 * nbr[q] + 1] - x[1], r2 = xs[3 * nbr[q] + 2] - x[2];
    float cur = sqrtf(fmaxf(r0 * r0 + r1 * r1 + r2 * r2, 1e-12f));
    float e = cur - L0[q];
    // ((stiffness * rel) / cur * (cur - L0)) / L0, evaluated left to right as :266-267 does, with the six
    // divisions per link sharing two reciprocals (48 IEEE divisions per node and substep dominated the kernel)
    const float ic = __frcp_rn(cur);
    fs[0] += div_rn(div_rn(stiff * r0, cur, ic) * e, L0[q], iL0[q]);
    fs[1] += div_rn(div_rn(stiff * r1, cur, ic) * e, L0[q], iL0[q]);
    fs[2] += div_rn(div_rn(stiff * r2, cur, ic) * e, L0[q], iL0[q]);
  }
  o.fr[0] = fs[0];
  o.fr[1] = fs[1] - k.g;
  o.fr[2] = fs[2];
  // Coulomb ground friction (:281-306)
  o.fmask = x[1] <= k.small_num;
  o.muF = mu * fminf(o.fr[1], 0.f) * -1.f;
  o.sV = sqrtf(o.vg[0] * o.vg[0] + o.vg[2] * o.vg[2] + k.small_num);
  o.dyn = o.fmask && (o.sV > k.small_num);
  float fx = o.fr[0], fz = o.fr[2];
  if (o.dyn) {
    const float is = __frcp_rn(o.sV);
    fx = fx - div_rn(o.muF * o.vg[0], o.sV, is);
    fz = fz - div_rn(o.muF * o.vg[2], o.sV, is);
  }
  o.stat = o.fmask && (o.sV <= k.small_num);
  o.sF = sqrtf(fx * fx + fz * fz + k.small_num);
  o.zero = o.stat && (o.muF > o.sF);
  o.nonz = o.stat && (o.muF <= o.sF);
  float xF = fx, yF = fz;
  if (o.zero) {
    fx = 0.f;
    fz = 0.f;
  }
  if (o.nonz) {
    float R = 1.f - o.muF / o.sF;
    fx = R * xF;
    fz = R * yF;
  }
  o.f[0] = fx;
  o.f[1] = o.fr[1];
  o.f[2] = fz;
#pragma unroll
  for (int c = 0; c < 3; ++c) o.v0[c] = (o.vg[c] + o.f[c] * k.dt) * k.damp;
  // grippers (:198-226)
  {
    float d0 = x[0] - ps0[0], d1 = x[1] - ps0[1], d2 = x[2] - ps0[2];
    o.m0 = sqrtf(d0 * d0 + d1 * d1 + d2 * d2) <= ps0[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o.v1[c] = o.m0 ? a0[3] * o.v0[c] : o.v0[c];
      o.x1[c] = o.m0 ? x[c] + a0[c] * (1.f - a0[3]) : x[c];
    }
    d0 = o.x1[0] - ps1[0];
    d1 = o.x1[1] - ps1[1];
    d2 = o.x1[2] - ps1[2];
    o.m1 = sqrtf(d0 * d0 + d1 * d1 + d2 * d2) <= ps1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o.v2[c] = o.m1 ? a1[3] * o.v1[c] : o.v1[c];
      o.x2[c] = o.m1 ? o.x1[c] + a1[c] * (1.f - a1[3]) : o.x1[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float vc = clipf(o.v2[c], -k.max_v, k.max_v);
    v[c] = vc;
    x[c] = clipf(o.x2[c], 0.f, 1.f) + k.dt * vc;
  }
}

__device__ __forceinline__ void advance_gripper(const float a[4], float ps[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) ps[c] = clipf(ps[c] + (c < 3 ? a[c] : 0.f), 0.f, 1.f);
}

__device__ __forceinline__ void load_actions(const float* __restrict__ action, int env, float a0[4], float a1[4]) {
  const float* a = action + (size_t)env * 8;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    a0[c] = clipf(a[c], -2.f, 2.f) / 50.f;       // :168
    a1[c] = clipf(a[4 + c], -2.f, 2.f) / 50.f;   // :169
  }
  a0[3] = a[3];
  a1[3] = a[7];
}

// save layout per (env, substep): [P*3 x][P*3 v][ps0 4][ps1 4]
__device__ __forceinline__ size_t save_stride(const ClothK& k) { return (size_t)k.P * 6 + 8; }

template <int MAXT, bool CLUSTER>
__global__ void __launch_bounds__(MAXT)
k_cloth_fwd(ClothK k, ud_cloth_state in, const int32_t* __restrict__ nbr_t, const float* __restrict__ L0_t,
            const float* __restrict__ action, ud_cloth_state out, float* __restrict__ save, int T,
            float* __restrict__ ckpt) {
  // T sub-actions (action is [T,B,8]) run back to back with the node state in registers; `ckpt` (nullable) receives
  // the state at the START of every sub-action: x [T,B,P,3] | v [T,B,P,3] | primitive0 [T,B,4] | primitive1 [T,B,4]
  extern __shared__ float xs[];  // [P*3]
  typedef Team<CLUSTER> TM;
  const int env = TM::env(), t = threadIdx.x;
  const int nn = TM::rank() * k.threads + t;   // node of this thread
  const bool live = nn < k.P;
  const bool lead = nn == 0;                   // the env's one writer of per-env values
  const int n = live ? nn : 0;
  const size_t o = (size_t)env * k.P + n;
  float x[3], v[3], L0[8], iL0[8], a0[4], a1[4], ps0[4], ps1[4];
  int nbr[8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    x[c] = in.x[3 * o + c];
    v[c] = in.v[3 * o + c];
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    nbr[q] = live ? nbr_t[n * 8 + q] : -1;
    L0[q] = L0_t[n * 8 + q];
    iL0[q] = __frcp_rn(L0[q]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ps0[c] = in.primitive0[env * 4 + c];
    ps1[c] = in.primitive1[env * 4 + c];
  }
  const float stiff = in.stiffness[env], mu = in.mu[env];
  for (int ta = 0; ta < T; ++ta) {
  load_actions(action + (size_t)ta * k.B * 8, env, a0, a1);
  if (ckpt) {
    const size_t BP3 = (size_t)k.B * k.P * 3;
    float* cx = ckpt + (size_t)ta * BP3 + 3 * o;
    float* cv = ckpt + (size_t)T * BP3 + (size_t)ta * BP3 + 3 * o;
    if (live) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        cx[c] = x[c];
        cv[c] = v[c];
      }
    }
    if (lead) {
      float* c0 = ckpt + 2 * (size_t)T * BP3 + ((size_t)ta * k.B + env) * 4;
      float* c1 = c0 + (size_t)T * k.B * 4;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        c0[c] = ps0[c];
        c1[c] = ps1[c];
      }
    }
  }
  for (int s = 0; s < k.S; ++s) {
    if (save) {
      float* sv = save + ((size_t)env * k.S + s) * save_stride(k);
      if (live) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          sv[3 * n + c] = x[c];
          sv[(size_t)k.P * 3 + 3 * n + c] = v[c];
        }
      }
      if (lead) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          sv[(size_t)k.P * 6 + c] = ps0[c];
          sv[(size_t)k.P * 6 + 4 + c] = ps1[c];
        }
      }
    }
    if (live) TM::publish(xs, n, x);
    TM::sync();
    Sub sb;
    cloth_substep(k, xs, nbr, L0, iL0, stiff, mu, a0, a1, ps0, ps1, x, v, sb);
    advance_gripper(a0, ps0);
    advance_gripper(a1, ps1);
    TM::sync();
  }
  }  // sub-actions
  if (!out.x) return;
  if (live) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      out.x[3 * o + c] = x[c];
      out.v[3 * o + c] = v[c];
    }
  }
  if (lead) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      out.primitive0[env * 4 + c] = ps0[c];
      out.primitive1[env * 4 + c] = ps1[c];
      out.action0[env * 4 + c] = a0[c];
      out.action1[env * 4 + c] = a1[c];
    }
    out.stiffness[env] = stiff;
    out.mu[env] = mu;
  }
}

// norm_grad backward for a 4-vector that every thread holds identically (gripper state)
__device__ __forceinline__ void norm_grad4(const ClothK& k, float g[4]) {
  float n2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2] + g[3] * g[3];
  float nrm = sqrtf(n2);
  const float rn = __frcp_rn(nrm), rms = __frcp_rn(k.mask_sum);
#pragma unroll
  for (int c = 0; c < 4; ++c) g[c] = norm_div(g[c], nrm, rn, k.mask_sum, rms);
}

// Adjoint of the 50-substep sub-action.  Expects `save` filled by k_cloth_fwd (recompute pass).
template <int MAXT, bool CLUSTER>
__global__ void __launch_bounds__(MAXT)
k_cloth_bwd(ClothK k, ud_cloth_state in, const int32_t* __restrict__ nbr_t, const float* __restrict__ L0_t,
            const float* __restrict__ action, ud_cloth_state gout, ud_cloth_state gin,
            float* __restrict__ gaction, const float* __restrict__ save) {
  extern __shared__ float sm[];
  typedef Team<CLUSTER> TM;
  float* xs = sm;                          // [P*3] positions of ALL nodes of the env (a copy per CTA)
  float* grel = sm + 3 * k.P;              // [threads*8*3] cotangents of the spring vectors of THIS CTA's nodes
  float* red = grel + 24 * k.threads;      // [64]
  float* redc = red + 64;                  // [2][16] cluster partials
  int par = 0;
  const int env = TM::env(), t = threadIdx.x;
  const int nn = TM::rank() * k.threads + t;
  const bool live = nn < k.P;
  const int n = live ? nn : 0;
  const size_t o = (size_t)env * k.P + n;
  const int mirror[8] = {1, 0, 3, 2, 7, 6, 5, 4};  // link k of i  <->  link mirror[k] of its neighbour
  float L0[8], iL0[8], a0[4], a1[4];
  int nbr[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    nbr[q] = live ? nbr_t[n * 8 + q] : -1;
    L0[q] = L0_t[n * 8 + q];
    iL0[q] = __frcp_rn(L0[q]);
  }
  load_actions(action, env, a0, a1);
  const float stiff = in.stiffness[env], mu = in.mu[env];
  const float rms = __frcp_rn(k.mask_sum);
  // incoming cotangents
  float gx[3], gv[3], gps0[4], gps1[4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    gx[c] = (live && gout.x) ? gout.x[3 * o + c] : 0.f;
    gv[c] = (live && gout.v) ? gout.v[3 * o + c] : 0.f;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    gps0[c] = gout.primitive0 ? gout.primitive0[env * 4 + c] : 0.f;
    gps1[c] = gout.primitive1 ? gout.primitive1[env * 4 + c] : 0.f;
  }
  // per-thread partial sums (reduced at the end) and uniform parts of the action cotangents
  float ga0p[4] = {0.f, 0.f, 0.f, 0.f}, ga1p[4] = {0.f, 0.f, 0.f, 0.f}, gstiffp = 0.f, gmup = 0.f;
  float ga0u[4], ga1u[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ga0u[c] = gout.action0 ? gout.action0[env * 4 + c] : 0.f;   // out.action0 = the scaled sub-action
    ga1u[c] = gout.action1 ? gout.action1[env * 4 + c] : 0.f;
  }
  for (int s = k.S - 1; s >= 0; --s) {
    const float* sv = save + ((size_t)env * k.S + s) * save_stride(k);
    float x[3], v[3], ps0[4], ps1[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      x[c] = sv[3 * n + c];
      v[c] = sv[(size_t)k.P * 3 + 3 * n + c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ps0[c] = sv[(size_t)k.P * 6 + c];
      ps1[c] = sv[(size_t)k.P * 6 + 4 + c];
    }
    const float xin[3] = {x[0], x[1], x[2]};
    TM::sync();
    if (live) TM::publish(xs, n, x);
    TM::sync();
    Sub f;
    cloth_substep(k, xs, nbr, L0, iL0, stiff, mu, a0, a1, ps0, ps1, x, v, f);  // x,v now hold the outputs (unused)
    // ---- (1) trailing norm_grads on x', v', ps0', ps1' (:331-334)
    float nx = live ? gx[0] * gx[0] + gx[1] * gx[1] + gx[2] * gx[2] : 0.f;
    float nv = live ? gv[0] * gv[0] + gv[1] * gv[1] + gv[2] * gv[2] : 0.f;
    team_sum2<CLUSTER>(nx, nv, red, redc, par);
    nx = sqrtf(nx);
    nv = sqrtf(nv);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      gx[c] = norm_div(gx[c], nx, __frcp_rn(nx), k.mask_sum, rms);
      gv[c] = norm_div(gv[c], nv, __frcp_rn(nv), k.mask_sum, rms);
    }
    norm_grad4(k, gps0);
    norm_grad4(k, gps1);
    // ---- (2) x' = clip(x2,0,1) + dt clip(v2) ; v' = clip(v2, +-max_v)
    float gx2[3], gv2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gvc = gv[c] + k.dt * gx[c];
      gv2[c] = clip_grad(f.v2[c], -k.max_v, k.max_v) * gvc;
      gx2[c] = clip_grad(f.x2[c], 0.f, 1.f) * gx[c];
    }
    // ---- (3) grippers advance: ps' = clip(ps + [a,0], 0, 1)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float u0 = ps0[c] + (c < 3 ? a0[c] : 0.f), u1 = ps1[c] + (c < 3 ? a1[c] : 0.f);
      gps0[c] = clip_grad(u0, 0.f, 1.f) * gps0[c];
      gps1[c] = clip_grad(u1, 0.f, 1.f) * gps1[c];
      if (c < 3) {
        ga0u[c] += gps0[c];
        ga1u[c] += gps1[c];
      }
    }
    // ---- (4) gripper 1: norm_grad(x2), norm_grad(v2), then the where()
    nx = live ? gx2[0] * gx2[0] + gx2[1] * gx2[1] + gx2[2] * gx2[2] : 0.f;
    nv = live ? gv2[0] * gv2[0] + gv2[1] * gv2[1] + gv2[2] * gv2[2] : 0.f;
    team_sum2<CLUSTER>(nx, nv, red, redc, par);
    nx = sqrtf(nx);
    nv = sqrtf(nv);
    float gx1[3], gv1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gxq = norm_div(gx2[c], nx, __frcp_rn(nx), k.mask_sum, rms), gvq = norm_div(gv2[c], nv, __frcp_rn(nv), k.mask_sum, rms);
      gx1[c] = gxq;
      if (f.m1 && live) {
        gv1[c] = a1[3] * gvq;
        ga1p[3] += gvq * f.v1[c] - gxq * a1[c];
        ga1p[c] += gxq * (1.f - a1[3]);
      } else {
        gv1[c] = gvq;
      }
    }
    // ---- (5) gripper 0
    nx = live ? gx1[0] * gx1[0] + gx1[1] * gx1[1] + gx1[2] * gx1[2] : 0.f;
    nv = live ? gv1[0] * gv1[0] + gv1[1] * gv1[1] + gv1[2] * gv1[2] : 0.f;
    team_sum2<CLUSTER>(nx, nv, red, redc, par);
    nx = sqrtf(nx);
    nv = sqrtf(nv);
    float gxs[3], gv0[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gxq = norm_div(gx1[c], nx, __frcp_rn(nx), k.mask_sum, rms), gvq = norm_div(gv1[c], nv, __frcp_rn(nv), k.mask_sum, rms);
      gxs[c] = gxq;  // x passes through either branch of the where()
      if (f.m0 && live) {
        gv0[c] = a0[3] * gvq;
        ga0p[3] += gvq * f.v0[c] - gxq * a0[c];
        ga0p[c] += gxq * (1.f - a0[3]);
      } else {
        gv0[c] = gvq;
      }
    }
    // ---- (6) v0 = (vg + f dt) damp
    float gf[3], gvg[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      gvg[c] = gv0[c] * k.damp;
      gf[c] = gv0[c] * k.damp * k.dt;
    }
    // ---- (7) friction reverse -> cotangent of the raw force fr, of vg (x,z) and of mu
    float gfr[3] = {gf[0], gf[1], gf[2]};
    float gmuF = 0.f;
    {
      // recompute the pre-static values
      float fx1 = f.fr[0], fz1 = f.fr[2];
      const float isV = __frcp_rn(f.sV);
      if (f.dyn) {
        fx1 = fx1 - div_rn(f.muF * f.vg[0], f.sV, isV);
        fz1 = fz1 - div_rn(f.muF * f.vg[2], f.sV, isV);
      }
      float gfx1 = gf[0], gfz1 = gf[2];
      if (f.zero) {
        gfx1 = 0.f;
        gfz1 = 0.f;
      }
      if (f.nonz) {
        // fx = R xF, fz = R yF, R = 1 - muF / sF, sF = sqrt(xF^2 + yF^2 + small)
        const float isF = __frcp_rn(f.sF);
        float R = 1.f - div_rn(f.muF, f.sF, isF);
        float gR = gf[0] * fx1 + gf[2] * fz1;
        gmuF += -gR * isF;
        float gsF = gR * f.muF * (isF * isF);
        gfx1 = gf[0] * R + gsF * fx1 * isF;
        gfz1 = gf[2] * R + gsF * fz1 * isF;
      }
      gfr[0] = gfx1;
      gfr[2] = gfz1;
      if (f.dyn) {
        // fx1 = fr0 - muF vg0 / sV ; sV = sqrt(vg0^2 + vg2^2 + small)
        const float gd = gfx1 * f.vg[0] + gfz1 * f.vg[2];
        gmuF += -gd * isV;
        float gsV = gd * f.muF * (isV * isV);
        gvg[0] += (-gfx1 * f.muF + gsV * f.vg[0]) * isV;
        gvg[2] += (-gfz1 * f.muF + gsV * f.vg[2]) * isV;
      }
      // muF = -mu * min(fr1, 0)
      if (live) gmup += gmuF * (-fminf(f.fr[1], 0.f));
      gfr[1] += (f.fr[1] < 0.f ? 1.f : (f.fr[1] == 0.f ? 0.5f : 0.f)) * gmuF * (-mu);
    }
    // ---- (8) spring forces: f_c = (stiff/L0) rel_c (1 - L0/cur)
    float gxi[3] = {gxs[0], gxs[1], gxs[2]};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (nbr[q] >= 0 && live) {
        float r0 = xs[3 * nbr[q]] - xin[0], r1 = xs[3 * nbr[q] + 1] - xin[1], r2 = xs[3 * nbr[q] + 2] - xin[2];
        float sq = r0 * r0 + r1 * r1 + r2 * r2;
        float cur = sqrtf(fmaxf(sq, 1e-12f));
        float e = cur - L0[q];
        float dot = gfr[0] * r0 + gfr[1] * r1 + gfr[2] * r2;
        const float inv = __frcp_rn(cur);
        float a = (stiff * iL0[q]) * (e * inv);                   // d f_c / d rel_c (direct)
        float b = sq >= 1e-12f ? stiff * dot * (inv * inv * inv) : 0.f;  // through cur
        g0 = a * gfr[0] + b * r0;
        g1 = a * gfr[1] + b * r1;
        g2 = a * gfr[2] + b * r2;
        if (k.stiff_float) gstiffp += dot * (e * inv) * iL0[q];
        gxi[0] -= g0;
        gxi[1] -= g1;
        gxi[2] -= g2;
      }
      if (live) {
        grel[(t * 8 + q) * 3] = g0;
        grel[(t * 8 + q) * 3 + 1] = g1;
        grel[(t * 8 + q) * 3 + 2] = g2;
      }
    }
    TM::sync();
    if (live) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (nbr[q] < 0) continue;
        // the neighbour's spring towards me, in the shared memory of the CTA that owns the neighbour
        const float* gown = grel;
        int nl = nbr[q];
        if (CLUSTER) {
          const int owner = nbr[q] / k.threads;
          nl = nbr[q] - owner * k.threads;
          gown = cg::this_cluster().map_shared_rank(grel, owner);
        }
        const float* gr = gown + (nl * 8 + mirror[q]) * 3;
        gxi[0] += gr[0];
        gxi[1] += gr[1];
        gxi[2] += gr[2];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      gx[c] = gxi[c];
      gv[c] = gvg[c];
    }
  }
  // ---- outputs
  if (live) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (gin.x) gin.x[3 * o + c] = gx[c];
      if (gin.v) gin.v[3 * o + c] = gv[c];
    }
  }
  // reduce the per-thread partials: (ga0p[0..3], ga1p[0..3], gstiffp, gmup)
  float pr[10] = {ga0p[0], ga0p[1], ga0p[2], ga0p[3], ga1p[0], ga1p[1], ga1p[2], ga1p[3], gstiffp, gmup};
  for (int i = 0; i < 10; i += 2) team_sum2<CLUSTER>(pr[i], pr[i + 1], red, redc, par);
  if (CLUSTER) TM::sync();   // no CTA leaves while its shared memory may still be read remotely
  if (nn == 0) {
    float ga0[4], ga1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ga0[c] = ga0u[c] + pr[c];
      ga1[c] = ga1u[c] + pr[4 + c];
      if (gin.primitive0) gin.primitive0[env * 4 + c] = gps0[c];
      if (gin.primitive1) gin.primitive1[env * 4 + c] = gps1[c];
      if (gin.action0) gin.action0[env * 4 + c] = 0.f;  // overwritten by robot_step (:173)
      if (gin.action1) gin.action1[env * 4 + c] = 0.f;
    }
    if (gin.stiffness) gin.stiffness[env] = pr[8] + (gout.stiffness ? gout.stiffness[env] : 0.f);
    if (gin.mu) gin.mu[env] = pr[9] + (gout.mu ? gout.mu[env] : 0.f);
    if (gaction) {
      const float* a = action + (size_t)env * 8;
      float* ga = gaction + (size_t)env * 8;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ga[c] = clip_grad(a[c], -2.f, 2.f) * (ga0[c] / 50.f);
        ga[4 + c] = clip_grad(a[4 + c], -2.f, 2.f) * (ga1[c] / 50.f);
      }
      ga[3] = ga0[3];
      ga[7] = ga1[3];
    }
  }
}

static bool cloth_state_ok(const ud_cloth_state* s) {
  return s && s->x && s->v && s->primitive0 && s->primitive1 && s->action0 && s->action1 && s->stiffness && s->mu;
}

// ---- launch helpers: one CTA per env up to 1024 nodes, a cluster of k.CL CTAs per env beyond --------------------
static size_t cloth_fwd_smem(const ClothK& k) { return sizeof(float) * 3 * (size_t)k.P; }
static size_t cloth_bwd_smem(const ClothK& k) { return sizeof(float) * (3 * (size_t)k.P + 24 * (size_t)k.threads + 64 + 32); }

template <class Kern, class... Args>
static void launch_cluster(Kern kern, const ClothK& k, size_t smem, cudaStream_t st, Args... args) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(k.B * k.CL));
  cfg.blockDim = dim3((unsigned)k.threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)k.CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, args...);
}

static void launch_cloth_fwd(const ClothK& k, cudaStream_t st, const ud_cloth_state& in, const int32_t* nbr, const float* L0,
                             const float* action, const ud_cloth_state& out, float* save, int T, float* ckpt) {
  const size_t smem = cloth_fwd_smem(k);
  if (k.CL > 1 && k.threads <= 512)
    launch_cluster(k_cloth_fwd<512, true>, k, smem, st, k, in, nbr, L0, action, out, save, T, ckpt);
  else if (k.CL > 1)
    launch_cluster(k_cloth_fwd<1024, true>, k, smem, st, k, in, nbr, L0, action, out, save, T, ckpt);
  else if (k.threads <= 512)
    k_cloth_fwd<512, false><<<k.B, k.threads, smem, st>>>(k, in, nbr, L0, action, out, save, T, ckpt);
  else
    k_cloth_fwd<1024, false><<<k.B, k.threads, smem, st>>>(k, in, nbr, L0, action, out, save, T, ckpt);
}

static void launch_cloth_bwd(const ClothK& k, cudaStream_t st, const ud_cloth_state& in, const int32_t* nbr, const float* L0,
                             const float* action, const ud_cloth_state& gout, const ud_cloth_state& gin, float* gaction,
                             const float* save) {
  const size_t smem = cloth_bwd_smem(k);
  if (k.CL > 1) {
    if (k.threads <= 512)
      launch_cluster(k_cloth_bwd<512, true>, k, smem, st, k, in, nbr, L0, action, gout, gin, gaction, save);
    else
      launch_cluster(k_cloth_bwd<1024, true>, k, smem, st, k, in, nbr, L0, action, gout, gin, gaction, save);
    return;
  }
  // per-DEVICE attribute (one host thread per device under pmap): set on every launch, like launch_cluster
  if (k.threads <= 512) {
    cudaFuncSetAttribute(k_cloth_bwd<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_cloth_bwd<512, false><<<k.B, k.threads, smem, st>>>(k, in, nbr, L0, action, gout, gin, gaction, save);
  } else {
    cudaFuncSetAttribute(k_cloth_bwd<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_cloth_bwd<1024, false><<<k.B, k.threads, smem, st>>>(k, in, nbr, L0, action, gout, gin, gaction, save);
  }
}

}  // namespace ud

using namespace ud;

extern "C" {

size_t ud_cloth_workspace_bytes(const ud_cloth_params* p) {
  ClothK k;
  if (!cloth_fold(p, &k)) return 0;
  return (((size_t)k.B * k.S * ((size_t)k.P * 6 + 8) * sizeof(float)) + 255) & ~(size_t)255;
}

int ud_cloth_step_fwd(const ud_cloth_params* p, const ud_cloth_state* in, const int32_t* nbr, const float* L0,
                      const float* action, ud_cloth_state* out, void* workspace, size_t workspace_bytes,
                      void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  ClothK k;
  if (!cloth_fold(p, &k)) return set_error(UD_E_INVALID, "ud_cloth_step_fwd: invalid ud_cloth_params (num_envs / n_nodes / substeps / T < 1, n_nodes above 8 CTAs x 1024 nodes, dt or mask_sum <= 0)");
  if (!cloth_state_ok(in) || !cloth_state_ok(out) || !nbr || !L0 || !action) return set_error(UD_E_INVALID, "ud_cloth_step_fwd: invalid argument (null pointer, size or parameter out of range)");
  cudaStream_t st = (cudaStream_t)stream;
  KScope ks(KC_CLOTH_FWD, st);
  launch_cloth_fwd(k, st, *in, nbr, L0, action, *out, nullptr, 1, nullptr);
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_cloth_step_fwd: launch failed");
}

int ud_cloth_step_bwd(const ud_cloth_params* p, const ud_cloth_state* in, const int32_t* nbr, const float* L0,
                      const float* action, const ud_cloth_state* gout, ud_cloth_state* gin, float* gaction,
                      void* workspace, size_t workspace_bytes, void* stream) {
  ClothK k;
  if (!cloth_fold(p, &k)) return set_error(UD_E_INVALID, "ud_cloth_step_bwd: invalid ud_cloth_params (num_envs / n_nodes / substeps / T < 1, n_nodes above 8 CTAs x 1024 nodes, dt or mask_sum <= 0)");
  if (!cloth_state_ok(in) || !gout || !gin || !nbr || !L0 || !action) return set_error(UD_E_INVALID, "ud_cloth_step_bwd: invalid argument (null pointer, size or parameter out of range)");
  size_t need = ud_cloth_workspace_bytes(p);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) return set_error(UD_E_WORKSPACE, "ud_cloth_step_bwd: workspace / checkpoint buffer too small or not 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  ud_cloth_state none;
  memset(&none, 0, sizeof(none));
  {
    KScope ks(KC_CLOTH_FWD, st);  // recompute pass: checkpoint = the sub-action input
    launch_cloth_fwd(k, st, *in, nbr, L0, action, none, (float*)workspace, 1, nullptr);
  }
  {
    KScope ks(KC_CLOTH_BWD, st);
    launch_cloth_bwd(k, st, *in, nbr, L0, action, *gout, *gin, gaction, (const float*)workspace);
  }
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_cloth_step_bwd: launch failed");
}

// ---- fused env step: T sub-actions per call (cloth_env.py:211 scans robot_step over the 40 pick-and-place
// sub-actions).  Forward = ONE launch for T*substeps substeps; the adjoint walks the sub-actions backwards from the
// checkpoints the forward left (recompute 50 substeps, reverse them), all enqueued from here: 2T launches, no host
// round trips.
static size_t cloth_ckpt_floats(const ClothK& k, int T) { return (size_t)T * k.B * ((size_t)k.P * 6 + 8); }
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t ud_cloth_multi_ckpt_bytes(const ud_cloth_params* p, int32_t T) {
  ClothK k;
  if (!cloth_fold(p, &k) || T < 1) return 0;
  return al256(cloth_ckpt_floats(k, T) * sizeof(float));
}

size_t ud_cloth_multi_workspace_bytes(const ud_cloth_params* p, int32_t T) {
  ClothK k;
  if (!cloth_fold(p, &k) || T < 1) return 0;
  // per-substep save slots of one sub-action + two cotangent sets (x, v, primitive0/1, action0/1, stiffness, mu)
  return ud_cloth_workspace_bytes(p) + 2 * al256(sizeof(float) * (size_t)k.B * ((size_t)k.P * 6 + 18));
}

int ud_cloth_multi_step_fwd(const ud_cloth_params* p, const ud_cloth_state* in, const int32_t* nbr, const float* L0,
                            const float* actions, int32_t T, ud_cloth_state* out, void* ckpt, size_t ckpt_bytes,
                            void* stream) {
  ClothK k;
  if (!cloth_fold(p, &k) || T < 1) return set_error(UD_E_INVALID, "ud_cloth_multi_step_fwd: invalid ud_cloth_params (num_envs / n_nodes / substeps / T < 1, n_nodes above 8 CTAs x 1024 nodes, dt or mask_sum <= 0)");
  if (!cloth_state_ok(in) || !cloth_state_ok(out) || !nbr || !L0 || !actions) return set_error(UD_E_INVALID, "ud_cloth_multi_step_fwd: invalid argument (null pointer, size or parameter out of range)");
  if (ckpt && (ckpt_bytes < ud_cloth_multi_ckpt_bytes(p, T) || ((uintptr_t)ckpt & 255))) return set_error(UD_E_WORKSPACE, "ud_cloth_multi_step_fwd: workspace / checkpoint buffer too small or not 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  KScope ks(KC_CLOTH_FWD, st);
  launch_cloth_fwd(k, st, *in, nbr, L0, actions, *out, nullptr, T, (float*)ckpt);
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_cloth_multi_step_fwd: launch failed");
}

int ud_cloth_multi_step_bwd(const ud_cloth_params* p, const ud_cloth_state* in, const int32_t* nbr, const float* L0,
                            const float* actions, int32_t T, const void* ckpt, const ud_cloth_state* gout,
                            ud_cloth_state* gin, float* gactions, void* workspace, size_t workspace_bytes,
                            void* stream) {
  ClothK k;
  if (!cloth_fold(p, &k) || T < 1) return set_error(UD_E_INVALID, "ud_cloth_multi_step_bwd: invalid ud_cloth_params (num_envs / n_nodes / substeps / T < 1, n_nodes above 8 CTAs x 1024 nodes, dt or mask_sum <= 0)");
  if (!cloth_state_ok(in) || !gout || !gin || !nbr || !L0 || !actions || !ckpt || !gactions) return set_error(UD_E_INVALID, "ud_cloth_multi_step_bwd: invalid argument (null pointer, size or parameter out of range)");
  if (!gin->x || !gin->v || !gin->primitive0 || !gin->primitive1 || !gin->stiffness || !gin->mu) return set_error(UD_E_INVALID, "ud_cloth_multi_step_bwd: invalid argument (null pointer, size or parameter out of range)");
  if (!workspace || workspace_bytes < ud_cloth_multi_workspace_bytes(p, T) || ((uintptr_t)workspace & 255))
    return set_error(UD_E_WORKSPACE, "ud_cloth_multi_step_bwd: workspace / checkpoint buffer too small or not 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t BP3 = (size_t)k.B * k.P * 3;
  float* save = (float*)workspace;
  // two cotangent sets for the ping-pong between sub-actions
  ud_cloth_state g[2];
  char* base = (char*)workspace + ud_cloth_workspace_bytes(p);
  for (int i = 0; i < 2; ++i) {
    float* f = (float*)(base + i * al256(sizeof(float) * (size_t)k.B * ((size_t)k.P * 6 + 18)));
    g[i].x = f;
    g[i].v = f + BP3;
    g[i].primitive0 = f + 2 * BP3;
    g[i].primitive1 = g[i].primitive0 + 4 * k.B;
    g[i].action0 = g[i].primitive1 + 4 * k.B;
    g[i].action1 = g[i].action0 + 4 * k.B;
    g[i].stiffness = g[i].action1 + 4 * k.B;
    g[i].mu = g[i].stiffness + k.B;
  }
  ud_cloth_state none;
  memset(&none, 0, sizeof(none));
  const float* ck = (const float*)ckpt;
  for (int t = T - 1; t >= 0; --t) {
    ud_cloth_state s_in = *in;  // state at the start of sub-action t (stiffness / mu are constants of the call)
    s_in.x = const_cast<float*>(ck + (size_t)t * BP3);
    s_in.v = const_cast<float*>(ck + (size_t)T * BP3 + (size_t)t * BP3);
    s_in.primitive0 = const_cast<float*>(ck + 2 * (size_t)T * BP3 + (size_t)t * k.B * 4);
    s_in.primitive1 = s_in.primitive0 + (size_t)T * k.B * 4;
    const float* a_t = actions + (size_t)t * k.B * 8;
    const ud_cloth_state& go = (t == T - 1) ? *gout : g[(t + 1) & 1];
    const ud_cloth_state& gi = (t == 0) ? *gin : g[t & 1];
    {
      KScope ks(KC_CLOTH_FWD, st);
      launch_cloth_fwd(k, st, s_in, nbr, L0, a_t, none, save, 1, nullptr);
    }
    {
      KScope ks(KC_CLOTH_BWD, st);
      launch_cloth_bwd(k, st, s_in, nbr, L0, a_t, go, gi, gactions + (size_t)t * k.B * 8, save);
    }
  }
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_cloth_multi_step_bwd: launch failed");
}

}  // extern "C"
