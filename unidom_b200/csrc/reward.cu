// Reward kernels either side of the simulator step inside the differentiated rollout:
// calc_chamfer (DaXBench/daxbench/core/utils/util.py:138-153) and calc_l2 (:156-159) with their adjoints.
//
// The reference materialises the (B,P,Q,3) difference tensor; here the nearest-neighbour scans keep one point per
// thread in registers and stream the other set through shared memory, so HBM sees x, y and 8 B per point of residuals
// (min squared distance + number of ties).  The distance of the reference is sqrt(mean_c (a-b)^2), monotone in the
// squared distance d2 = sum_c (a-b)^2, so the scans work on d2 and take the square root once per point.
// jnp.min's VJP splits the cotangent evenly over exact ties (lax reduce chooser rule): the tie count is part of the
// residuals and the adjoint re-evaluates d2 with the same rounding (explicit _rn intrinsics, no FMA contraction).
#include "mpm_internal.h"

namespace ud {

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
constexpr int NN_BLOCK = 256;
constexpr int NN_TILE = 1024;

__device__ __forceinline__ float d2_rn(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// For every point a of set A (per env when strideA != 0, shared otherwise): min over set B of d2, and how many
// points of B attain it.  NaN anywhere in a row makes the row's minimum NaN (jnp.min propagates NaN).
__global__ void __launch_bounds__(NN_BLOCK)
k_nn_scan(const float* __restrict__ A, size_t strideA, int nA, const float* __restrict__ Bp, size_t strideB, int nB,
          float* __restrict__ mind2, int32_t* __restrict__ cnt) {
  __shared__ float4 tile[NN_TILE];
  const int env = blockIdx.y;
  const int a = blockIdx.x * NN_BLOCK + threadIdx.x;
  const float* Ae = A + strideA * env;
  const float* Be = Bp + strideB * env;
  float ax = 0.f, ay = 0.f, az = 0.f;
  if (a < nA) { ax = Ae[3 * a]; ay = Ae[3 * a + 1]; az = Ae[3 * a + 2]; }
  float m = __int_as_float(0x7f800000);
  int c = 0;
  bool bad = false;
  for (int b0 = 0; b0 < nB; b0 += NN_TILE) {
    const int nt = min(NN_TILE, nB - b0);
    __syncthreads();
    for (int j = threadIdx.x; j < nt; j += NN_BLOCK)
      tile[j] = make_float4(Be[3 * (b0 + j)], Be[3 * (b0 + j) + 1], Be[3 * (b0 + j) + 2], 0.f);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < nt; ++j) {
      const float4 q = tile[j];                       // same address for the whole warp: one broadcast LDS.128
      const float d = d2_rn(ax, ay, az, q.x, q.y, q.z);
      bad |= !(d == d);
      if (d < m) { m = d; c = 1; } else if (d == m) ++c;
    }
  }
  if (a < nA) {
    mind2[(size_t)env * nA + a] = bad ? __int_as_float(0x7fc00000) : m;
    cnt[(size_t)env * nA + a] = c;
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = l < (int)(blockDim.x >> 5) ? red[l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;      // valid in warp 0
}

// out[env] = mean_q sqrt(miny/3) + mean_p sqrt(minx/3)   (util.py:151: y2x_min + x2y_min); one CTA per env,
// fixed summation order.
__global__ void __launch_bounds__(256)
k_chamfer_finish(const float* __restrict__ minx, int P, const float* __restrict__ miny, int Q, float* __restrict__ out) {
  __shared__ float red[8];
  const int env = blockIdx.x;
  float sx = 0.f, sy = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) sx += sqrtf(minx[(size_t)env * P + p] / 3.f);
  for (int q = threadIdx.x; q < Q; q += blockDim.x) sy += sqrtf(miny[(size_t)env * Q + q] / 3.f);
  sx = block_sum(sx, red);
  sy = block_sum(sy, red);
  if (threadIdx.x == 0) out[env] = sy / (float)Q + sx / (float)P;
}

// d chamfer / d x.  Thread = one point x_p; streams y together with the y-side residuals.
//   x2y term: g/P * (x_p - y_q) / (3 sqrt(minx_p/3)) / cntx_p          for every q with d2(p,q) == minx_p
//   y2x term: g/Q * (x_p - y_q) / (3 sqrt(miny_q/3)) / cnty_q          for every q with d2(p,q) == miny_q
// (a point lying exactly on its neighbour gives 0 * inf = NaN, as in the reference; APG scrubs it, apg.py:233).
__global__ void __launch_bounds__(NN_BLOCK)
k_chamfer_bwd(const float* __restrict__ x, const float* __restrict__ y, int P, int Q, const float* __restrict__ minx,
              const int32_t* __restrict__ cntx, const float* __restrict__ miny, const int32_t* __restrict__ cnty,
              const float* __restrict__ gout, float* __restrict__ gx) {
  __shared__ float4 tile[NN_TILE];
  __shared__ float coef[NN_TILE];
  const int env = blockIdx.y;
  const int p = blockIdx.x * NN_BLOCK + threadIdx.x;
  const float g = gout[env];
  float ax = 0.f, ay = 0.f, az = 0.f, mp = -1.f;
  if (p < P) {
    const float* xe = x + ((size_t)env * P + p) * 3;
    ax = xe[0]; ay = xe[1]; az = xe[2];
    mp = minx[(size_t)env * P + p];
  }
  float sx[3] = {0.f, 0.f, 0.f}, sy[3] = {0.f, 0.f, 0.f};
  for (int b0 = 0; b0 < Q; b0 += NN_TILE) {
    const int nt = min(NN_TILE, Q - b0);
    __syncthreads();
    for (int j = threadIdx.x; j < nt; j += NN_BLOCK) {
      const float mq = miny[(size_t)env * Q + b0 + j];
      tile[j] = make_float4(y[3 * (b0 + j)], y[3 * (b0 + j) + 1], y[3 * (b0 + j) + 2], mq);
      coef[j] = g / ((float)Q * 3.f * sqrtf(mq / 3.f) * (float)cnty[(size_t)env * Q + b0 + j]);
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < nt; ++j) {
      const float4 q = tile[j];
      const float d = d2_rn(ax, ay, az, q.x, q.y, q.z);
      if (d == mp) { sx[0] += ax - q.x; sx[1] += ay - q.y; sx[2] += az - q.z; }
      if (d == q.w) {
        const float cf = coef[j];
        sy[0] += cf * (ax - q.x); sy[1] += cf * (ay - q.y); sy[2] += cf * (az - q.z);
      }
    }
  }
  if (p < P) {
    const float cx = g / ((float)P * 3.f * sqrtf(mp / 3.f) * (float)cntx[(size_t)env * P + p]);
    float* o = gx + ((size_t)env * P + p) * 3;
    const bool nanrow = !(mp == mp);               // NaN row: jnp.min's chooser matches nothing -> 0/0
    for (int c = 0; c < 3; ++c) o[c] = nanrow ? mp : (cntx[(size_t)env * P + p] ? cx * sx[c] : 0.f) + sy[c];
  }
}

// calc_l2 (util.py:156-159): out[env] = mean_p sqrt(mean_c (x - y)^2), one CTA per env.
__global__ void __launch_bounds__(256)
k_l2_fwd(const float* __restrict__ x, const float* __restrict__ y, int P, float* __restrict__ out) {
  __shared__ float red[8];
  const int env = blockIdx.x;
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float* xe = x + ((size_t)env * P + p) * 3;
    s += sqrtf(d2_rn(xe[0], xe[1], xe[2], y[3 * p], y[3 * p + 1], y[3 * p + 2]) / 3.f);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[env] = s / (float)P;
}

__global__ void k_l2_bwd(const float* __restrict__ x, const float* __restrict__ y, int B, int P,
                         const float* __restrict__ gout, float* __restrict__ gx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * P) return;
  const int env = (int)(i / P), p = (int)(i % P);
  const float dx = x[3 * i] - y[3 * p], dy = x[3 * i + 1] - y[3 * p + 1], dz = x[3 * i + 2] - y[3 * p + 2];
  const float u = d2_rn(x[3 * i], x[3 * i + 1], x[3 * i + 2], y[3 * p], y[3 * p + 1], y[3 * p + 2]) / 3.f;
  const float cf = gout[env] / ((float)P * 3.f * sqrtf(u));
  gx[3 * i] = cf * dx;
  gx[3 * i + 1] = cf * dy;
  gx[3 * i + 2] = cf * dz;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct ChamferRes { float* minx; int32_t* cntx; float* miny; int32_t* cnty; size_t bytes; };
static ChamferRes chamfer_carve(void* base, int B, int P, int Q) {
  char* b = (char*)base;
  size_t off = 0;
  ChamferRes r;
  r.minx = (float*)(b + off); off += al256(4 * (size_t)B * P);
  r.cntx = (int32_t*)(b + off); off += al256(4 * (size_t)B * P);
  r.miny = (float*)(b + off); off += al256(4 * (size_t)B * Q);
  r.cnty = (int32_t*)(b + off); off += al256(4 * (size_t)B * Q);
  r.bytes = off;
  return r;
}

}  // namespace ud

using namespace ud;

extern "C" {

size_t ud_chamfer_residual_bytes(int32_t B, int32_t P, int32_t Q) {
  if (B < 1 || P < 1 || Q < 1) return 0;
  return chamfer_carve(nullptr, B, P, Q).bytes;
}

int ud_chamfer_fwd(const float* x, const float* y, int32_t B, int32_t P, int32_t Q, float* out, void* residuals,
                   size_t residual_bytes, void* stream) {
  if (!x || !y || !out || B < 1 || P < 1 || Q < 1 || B > 65535) return set_error(UD_E_INVALID, "ud_chamfer_fwd: invalid argument (null pointer, size or parameter out of range)");
  if (!residuals || ((uintptr_t)residuals & 255) || residual_bytes < chamfer_carve(nullptr, B, P, Q).bytes)
    return set_error(UD_E_WORKSPACE, "ud_chamfer_fwd: workspace / checkpoint buffer too small or not 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  ChamferRes r = chamfer_carve(residuals, B, P, Q);
  {
    KScope ks(KC_REWARD, st, 3);
    k_nn_scan<<<dim3(cdiv(P, NN_BLOCK), B), NN_BLOCK, 0, st>>>(x, (size_t)P * 3, P, y, 0, Q, r.minx, r.cntx);
    k_nn_scan<<<dim3(cdiv(Q, NN_BLOCK), B), NN_BLOCK, 0, st>>>(y, 0, Q, x, (size_t)P * 3, P, r.miny, r.cnty);
    k_chamfer_finish<<<B, 256, 0, st>>>(r.minx, P, r.miny, Q, out);
  }
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_chamfer_fwd: launch failed");
}

int ud_chamfer_bwd(const float* x, const float* y, int32_t B, int32_t P, int32_t Q, const float* gout,
                   const void* residuals, size_t residual_bytes, float* gx, void* stream) {
  if (!x || !y || !gout || !gx || B < 1 || P < 1 || Q < 1 || B > 65535) return set_error(UD_E_INVALID, "ud_chamfer_bwd: invalid argument (null pointer, size or parameter out of range)");
  if (!residuals || ((uintptr_t)residuals & 255) || residual_bytes < chamfer_carve(nullptr, B, P, Q).bytes)
    return set_error(UD_E_WORKSPACE, "ud_chamfer_bwd: workspace / checkpoint buffer too small or not 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  ChamferRes r = chamfer_carve((void*)residuals, B, P, Q);
  {
    KScope ks(KC_REWARD, st);
    k_chamfer_bwd<<<dim3(cdiv(P, NN_BLOCK), B), NN_BLOCK, 0, st>>>(x, y, P, Q, r.minx, r.cntx, r.miny, r.cnty, gout, gx);
  }
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_chamfer_bwd: launch failed");
}

int ud_l2_fwd(const float* x, const float* y, int32_t B, int32_t P, float* out, void* stream) {
  if (!x || !y || !out || B < 1 || P < 1) return set_error(UD_E_INVALID, "ud_l2_fwd: invalid argument (null pointer, size or parameter out of range)");
  cudaStream_t st = (cudaStream_t)stream;
  KScope ks(KC_REWARD, st);
  k_l2_fwd<<<B, 256, 0, st>>>(x, y, P, out);
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_l2_fwd: launch failed");
}

int ud_l2_bwd(const float* x, const float* y, int32_t B, int32_t P, const float* gout, float* gx, void* stream) {
  if (!x || !y || !gout || !gx || B < 1 || P < 1) return set_error(UD_E_INVALID, "ud_l2_bwd: invalid argument (null pointer, size or parameter out of range)");
  cudaStream_t st = (cudaStream_t)stream;
  KScope ks(KC_REWARD, st);
  const size_t n = (size_t)B * P;
  k_l2_bwd<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, B, P, gout, gx);
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_l2_bwd: launch failed");
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// APG update on the flat policy-gradient buffer (DaXBench/daxbench/algorithms/apg/apg.py:233-240,
// 260-267): per-rank nan_to_num + global norm, clip to max_grad_norm, [all-reduce by the caller], mean + optax.adam.
// Three element-wise passes over ~1 M floats instead of ~15 framework launches; the collective stays NCCL.
// ------------------------------------------------------------------------------------------------------------------
namespace ud {

__global__ void __launch_bounds__(256) k_apg_scrub_sumsq(float* __restrict__ g, long long n, float* __restrict__ sumsq) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = g[i];
    t = t == t ? fminf(fmaxf(t, -3.4028234663852886e38f), 3.4028234663852886e38f) : 0.f;   // jnp.nan_to_num
    g[i] = t;
    acc += t * t;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(sumsq, acc);
}

// g <- where(norm < max_norm, g, (g / norm) * max_norm)   (apg.py:264-266, evaluated as written)
__global__ void __launch_bounds__(256)
k_apg_clip(float* __restrict__ g, long long n, const float* __restrict__ sumsq, float max_norm) {
  const float norm = sqrtf(*sumsq);
  if (norm < max_norm) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    g[i] = (g[i] / norm) * max_norm;
}

// g' = g / world (pmean after the caller's all-reduce SUM), then optax.adam: m, v, bias corrections c1 = 1 - b1^t,
// c2 = 1 - b2^t, p -= lr * (m / c1) / (sqrt(v / c2) + eps)
__global__ void __launch_bounds__(256)
k_adam_step(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            float world, float lr, float b1, float b2, float omb1, float omb2, float eps, float c1, float c2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = world == 1.f ? g[i] : g[i] / world;
    // one rounding per operation, like the op-by-op host evaluation (no FMA contraction)
    const float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(omb1, gi));
    const float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(__fmul_rn(omb2, gi), gi));
    m[i] = mi;
    v[i] = vi;
    const float mhat = __fdiv_rn(mi, c1), vhat = __fdiv_rn(vi, c2);
    p[i] = __fsub_rn(p[i], __fdiv_rn(__fmul_rn(lr, mhat), __fadd_rn(__fsqrt_rn(vhat), eps)));
  }
}

}  // namespace ud

extern "C" {

int ud_apg_scrub_clip(float* grad, int64_t n, float max_grad_norm, float* sumsq, void* stream) {
  if (!grad || !sumsq || n < 1) return set_error(UD_E_INVALID, "ud_apg_scrub_clip: invalid argument (null pointer, size or parameter out of range)");
  cudaStream_t st = (cudaStream_t)stream;
  KScope ks(KC_REWARD, st, 2);
  cudaMemsetAsync(sumsq, 0, sizeof(float), st);
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_apg_scrub_sumsq<<<blocks, 256, 0, st>>>(grad, n, sumsq);
  k_apg_clip<<<blocks, 256, 0, st>>>(grad, n, sumsq, max_grad_norm);
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_apg_scrub_clip: launch failed");
}

int ud_adam_step(float* params, const float* grad, float* m, float* v, int64_t n, int32_t world_size, double lr, double b1,
                 double b2, double eps, int32_t t, void* stream) {
  if (!params || !grad || !m || !v || n < 1 || world_size < 1 || t < 1) return set_error(UD_E_INVALID, "ud_adam_step: invalid argument (null pointer, size or parameter out of range)");
  cudaStream_t st = (cudaStream_t)stream;
  KScope ks(KC_REWARD, st);
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  // scalars evaluated in double and rounded once, as Python evaluates optax's / apg.Adam's
  const float c1 = (float)(1.0 - pow(b1, (double)t)), c2 = (float)(1.0 - pow(b2, (double)t));
  k_adam_step<<<blocks, 256, 0, st>>>(params, grad, m, v, n, (float)world_size, (float)lr, (float)b1, (float)b2,
                                      (float)(1.0 - b1), (float)(1.0 - b2), (float)eps, c1, c2);
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_adam_step: launch failed");
}

}  // extern "C"
