// APG update as ONE kernel per rank: scrub + per-rank global-norm clip -> mean over ranks -> Adam
// (DaXBench/daxbench/algorithms/apg/apg.py:233-240 `clip_by_global_norm` + `pmean`, :260-267 optax.adam).
//
// The three-kernel path (ud_apg_scrub_clip, NCCL all-reduce, ud_adam_step) moves the 3.7 MB gradient through HBM five
// times and costs four launches plus NCCL's own; here the collective is part of the kernel.  Every rank
//   1. scrubs its gradient and reduces its squared norm (grid barrier),
//   2. stages the clipped gradient in a buffer that is mapped into every peer (NVLink peer memory; the caller obtains
//      the mapping, e.g. torch symmetric memory / cudaIpc / cuMem fabric handles -- plain device pointers here),
//   3. publishes "iteration t staged" in every peer's flag array (st.release.sys) and waits for the peers' flags,
//   4. reads ALL ranks' staged gradients with P2P loads in rank order 0..N-1 -- the same order on every rank, so the
//      replicas stay bit-identical without a broadcast --, divides by N and applies Adam to its replica.
// Staging is double-buffered by iteration parity: a rank may stage iteration t+1 while a peer still reads t; it can
// only reach t+2 after every peer has signalled t+1, i.e. has finished reading t.
// From UD_APG_RS_MIN_WORLD ranks up (ud_tuning_set("apg_rs", 0 / 1) forces either form) step 4 is reduce-scatter +
// broadcast (RS): a rank sums only ITS slice of the staged
// gradients (rank order again), stores the mean of the slice into every peer's `reduced` buffer (P2P stores), the ranks
// meet at a second flag barrier and Adam then reads local memory only: 2 n instead of N n elements over NVLink per rank
// (at N = 8 the scalar all-read form tied with NCCL: 124 vs 118 us per update; at N = 2 the two forms move the same bytes
// and the all-read form wins by its one barrier less, 48 vs 54 us).  Same sums in the same order as the all-read form.
// Arithmetic and rounding are those of k_apg_clip / k_adam_step (csrc/reward.cu): op-by-op, no FMA contraction.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/unidom_b200.h"
#include "mpm_internal.h"

namespace ud {

namespace {

constexpr int FB = 256;   // threads per CTA

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all CTAs of the grid are resident (the launch is sized for that): arrive + spin
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_gpu(ctr) < nblocks) {
    }
  }
  __syncthreads();
}
__device__ __forceinline__ float scrub(float t) {   // jnp.nan_to_num
  return t == t ? fminf(fmaxf(t, -3.4028234663852886e38f), 3.4028234663852886e38f) : 0.f;
}

// scratch: [0] sum of squares (float), [1..3] grid-barrier counters (unsigned); zeroed by the launcher
template <bool RS>
__global__ void __launch_bounds__(FB)
k_apg_fused(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v, long long n,
            float max_norm, float lr, float b1, float b2, float omb1, float omb2, float eps, float c1, float c2, int t,
            int rank, int world, const unsigned long long* __restrict__ peer_stage,
            const unsigned long long* __restrict__ peer_flags, float* __restrict__ scratch, bool vec) {
  __shared__ float red[FB / 32];
  unsigned* ctr = reinterpret_cast<unsigned*>(scratch) + 1;
  const long long tid = (long long)blockIdx.x * FB + threadIdx.x, nthr = (long long)gridDim.x * FB;
  const size_t slot = (size_t)(t & 1) * (size_t)n;
  float* my_stage = reinterpret_cast<float*>(peer_stage[rank]) + slot;
  // ---- 1. per-rank squared norm of the scrubbed gradient
  // vec (launcher: n % 4 == 0 and 16-byte aligned arrays): every pass moves float4; per element the same operations
  const long long n4 = n >> 2;
  float acc = 0.f;
  if (vec) {
    for (long long q = tid; q < n4; q += nthr) {
      const float4 g4 = reinterpret_cast<const float4*>(grad)[q];
      const float g0 = scrub(g4.x), g1 = scrub(g4.y), g2 = scrub(g4.z), g3 = scrub(g4.w);
      acc += g0 * g0;
      acc += g1 * g1;
      acc += g2 * g2;
      acc += g3 * g3;
    }
  } else {
    for (long long i = tid; i < n; i += nthr) {
      const float g = scrub(grad[i]);
      acc += g * g;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < FB / 32; ++i) s += red[i];
    atomicAdd(scratch, s);
  }
  grid_barrier(ctr + 0, gridDim.x);
  // ---- 2. clip (apg.py:264-266, evaluated as written) and stage for the peers
  const float norm = sqrtf(__ldcg(scratch));
  const bool clip = !(norm < max_norm);
  if (vec) {
    for (long long q = tid; q < n4; q += nthr) {
      const float4 g4 = reinterpret_cast<const float4*>(grad)[q];
      const float g0 = scrub(g4.x), g1 = scrub(g4.y), g2 = scrub(g4.z), g3 = scrub(g4.w);
      reinterpret_cast<float4*>(my_stage)[q] =
          clip ? make_float4((g0 / norm) * max_norm, (g1 / norm) * max_norm, (g2 / norm) * max_norm, (g3 / norm) * max_norm)
               : make_float4(g0, g1, g2, g3);
    }
  } else {
    for (long long i = tid; i < n; i += nthr) {
      const float g = scrub(grad[i]);
      my_stage[i] = clip ? (g / norm) * max_norm : g;
    }
  }
  __threadfence_system();
  grid_barrier(ctr + 1, gridDim.x);
  // ---- 3. flag barrier over the ranks in peer memory
  if (world > 1) {
    if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(reinterpret_cast<int*>(peer_flags[threadIdx.x]) + rank, t);
    if (threadIdx.x < world) {
      const int* mine = reinterpret_cast<const int*>(peer_flags[rank]) + threadIdx.x;
      while (ld_acquire_sys(mine) < t) {
      }
    }
    __syncthreads();
  }
  // ---- 4. mean over the ranks (rank order: identical on every replica) + optax.adam
  const float fw = (float)world;
  const float* reduced = nullptr;
  if (RS) {
    // my slice, in whole float4 (the launcher picks this form only when vec holds, so there is no tail)
    const long long per = (n4 + world - 1) / world;
    const long long q0 = per * rank < n4 ? per * rank : n4, q1 = q0 + per < n4 ? q0 + per : n4;
    const size_t rslot = (size_t)(2 + (t & 1)) * (size_t)n;   // `reduced` lives behind the two staging slots
    for (long long q = q0 + tid; q < q1; q += nthr) {
      float4 acc;
      for (int r = 0; r < world; ++r) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(peer_stage[r]) + slot) + q);
        if (r == 0) acc = a;
        else acc = make_float4(__fadd_rn(acc.x, a.x), __fadd_rn(acc.y, a.y), __fadd_rn(acc.z, a.z), __fadd_rn(acc.w, a.w));
      }
      acc = make_float4(acc.x / fw, acc.y / fw, acc.z / fw, acc.w / fw);
      for (int r = 0; r < world; ++r)
        __stcg(reinterpret_cast<float4*>(reinterpret_cast<float*>(peer_stage[r]) + rslot) + q, acc);
    }
    __threadfence_system();
    grid_barrier(ctr + 2, gridDim.x);
    // second flag barrier: every rank's slice has landed in my `reduced` buffer
    if (blockIdx.x == 0 && threadIdx.x < world)
      st_release_sys(reinterpret_cast<int*>(peer_flags[threadIdx.x]) + 32 + rank, t);
    if (threadIdx.x < world) {
      const int* mine = reinterpret_cast<const int*>(peer_flags[rank]) + 32 + threadIdx.x;
      while (ld_acquire_sys(mine) < t) {
      }
    }
    __syncthreads();
    reduced = reinterpret_cast<const float*>(peer_stage[rank]) + rslot;
  }
  auto adam = [&](float gi, float mo, float vo, float po, float& mn, float& vn, float& pn) {
    mn = __fadd_rn(__fmul_rn(b1, mo), __fmul_rn(omb1, gi));
    vn = __fadd_rn(__fmul_rn(b2, vo), __fmul_rn(__fmul_rn(omb2, gi), gi));
    const float mhat = __fdiv_rn(mn, c1), vhat = __fdiv_rn(vn, c2);
    pn = __fsub_rn(po, __fdiv_rn(__fmul_rn(lr, mhat), __fadd_rn(__fsqrt_rn(vhat), eps)));
  };
  if (vec) {
    for (long long q = tid; q < n4; q += nthr) {
      float4 g4;
      if (RS) {
        g4 = __ldcg(reinterpret_cast<const float4*>(reduced) + q);
      } else {
        for (int r = 0; r < world; ++r) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(peer_stage[r]) + slot) + q);
          if (r == 0) g4 = a;
          else g4 = make_float4(__fadd_rn(g4.x, a.x), __fadd_rn(g4.y, a.y), __fadd_rn(g4.z, a.z), __fadd_rn(g4.w, a.w));
        }
        if (world > 1) g4 = make_float4(g4.x / fw, g4.y / fw, g4.z / fw, g4.w / fw);
      }
      const float4 m4 = reinterpret_cast<const float4*>(m)[q], v4 = reinterpret_cast<const float4*>(v)[q],
                   p4 = reinterpret_cast<const float4*>(p)[q];
      float4 mn, vn, pn;
      adam(g4.x, m4.x, v4.x, p4.x, mn.x, vn.x, pn.x);
      adam(g4.y, m4.y, v4.y, p4.y, mn.y, vn.y, pn.y);
      adam(g4.z, m4.z, v4.z, p4.z, mn.z, vn.z, pn.z);
      adam(g4.w, m4.w, v4.w, p4.w, mn.w, vn.w, pn.w);
      reinterpret_cast<float4*>(m)[q] = mn;
      reinterpret_cast<float4*>(v)[q] = vn;
      reinterpret_cast<float4*>(p)[q] = pn;
    }
    return;
  }
  for (long long i = tid; i < n; i += nthr) {
    float gi;
    if (RS) {
      gi = __ldcg(reduced + i);
    } else {
      float s = 0.f;
      for (int r = 0; r < world; ++r) {
        const float* st = reinterpret_cast<const float*>(peer_stage[r]) + slot;
        s = r == 0 ? __ldcg(st + i) : __fadd_rn(s, __ldcg(st + i));   // .cg: peer lines are never cached in L1
      }
      gi = world == 1 ? s : s / fw;
    }
    float mi, vi, pi;
    adam(gi, m[i], v[i], p[i], mi, vi, pi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
  }
}

}  // namespace

// ud_tuning_set("apg_rs", v): -1 = reduce-scatter form from UD_APG_RS_MIN_WORLD ranks up (default), 0 = never, 1 = always
#ifndef UD_APG_RS_MIN_WORLD
#define UD_APG_RS_MIN_WORLD 4
#endif
static int g_apg_rs = -1;
int tuning_apg_rs(int v) {
  const int o = g_apg_rs;
  if (v >= -1 && v <= 1) g_apg_rs = v;
  return o;
}
}  // namespace ud

extern "C" int ud_apg_fused_update(float* params, const float* grad, float* m, float* v, int64_t n, float max_grad_norm,
                                   double lr, double b1, double b2, double eps, int32_t t, int32_t rank, int32_t world,
                                   const uint64_t* peer_stage, const uint64_t* peer_flags, float* scratch, void* stream) {
  using namespace ud;
  if (!params || !grad || !m || !v || !peer_stage || !peer_flags || !scratch || n < 1 || t < 1 || world < 1 || world > 32 ||
      rank < 0 || rank >= world)
    return set_error(UD_E_INVALID, "ud_apg_fused_update: invalid argument (null pointer, size, rank or world out of range)");
  cudaStream_t st = (cudaStream_t)stream;
  KScope ks(KC_APG, st);
  cudaMemsetAsync(scratch, 0, 8 * sizeof(float), st);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (n + FB - 1) / FB;
  const int blocks = (int)(want < 2LL * sms ? want : 2LL * sms);   // resident by construction: 2 x 256 threads per SM
  const float c1 = (float)(1.0 - pow(b1, (double)t)), c2 = (float)(1.0 - pow(b2, (double)t));
  const bool vec = n % 4 == 0 && (((uintptr_t)params | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0;   // (the
  // staging buffers are the caller's symmetric allocations: 16-byte aligned by contract, their slots by n % 4 == 0)
  const bool rs = vec && world > 1 && (g_apg_rs == 1 || (g_apg_rs < 0 && world >= UD_APG_RS_MIN_WORLD));
  auto kern = rs ? k_apg_fused<true> : k_apg_fused<false>;
  kern<<<blocks, FB, 0, st>>>(params, grad, m, v, (long long)n, max_grad_norm, (float)lr, (float)b1, (float)b2,
                                     (float)(1.0 - b1), (float)(1.0 - b2), (float)eps, c1, c2, (int)t, (int)rank, (int)world,
                                     reinterpret_cast<const unsigned long long*>(peer_stage),
                                     reinterpret_cast<const unsigned long long*>(peer_flags), scratch, vec);
  return cudaGetLastError() == cudaSuccess ? UD_OK : set_error(UD_E_CUDA, "ud_apg_fused_update: launch failed");
}
