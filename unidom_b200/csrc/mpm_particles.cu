// Particle-side kernels of the MLS-MPM step: per-frame binning, P2G scatter, G2P gather and their
// adjoints.  sm_100a.  Reference: DaXBench/daxbench/core/engine/mpm_simulator.py:178-330.
#include <algorithm>
#include <climits>
#include <type_traits>

#include "mpm_internal.h"

namespace ud {

#define UD_BLOCK 128

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }
// SM count of the CURRENT device (a process may drive several devices: no process-wide cache)
static inline int num_sms() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
static inline dim3 pgrid(const MpmConst& k, int block) { return dim3(cdiv(k.n, block), k.B); }

// particle index of this thread: blockIdx.y = env, blockIdx.x*blockDim.x + threadIdx.x = slot in env.
//   g  = env*n + slot      index into the per-particle arrays (perm, material, h, AoS leaves)
//   gp = env*n_pad + slot  index into the sorted tile arrays (mpm_internal.h); slots in [n, n_pad) are padding that
//                          exists in memory (dead lanes may read it, never write it)
#define UD_PARTICLE_INDEX(k, env, g)                         \
  int env = blockIdx.y;                                      \
  int slot_ = blockIdx.x * blockDim.x + threadIdx.x;         \
  bool live_ = slot_ < (k).n;                                \
  int g = env * (k).n + (live_ ? slot_ : 0);                 \
  int gp = env * (k).n_pad + (slot_ < (k).n_pad ? slot_ : 0);

// The same with the CTAs walking the tiles from the LAST one down (k_g2p): the persistent P2G of the next substep walks
// upwards and reads x, v, C of the tiles G2P wrote LAST first, while they are still in L2.  Measured on the plasticine
// scene: k_p2g_pers 123.2 -> 121.2 us per launch, k_g2p unchanged (profiles/r02_run25.sh).  The same reversal of G2P^T
// (after the upward P2G^T) cost it more than it gained: 110.9 -> 115.8 us, not kept.
#define UD_PARTICLE_INDEX_REV(k, env, g)                                              \
  int env = gridDim.y - 1 - blockIdx.y;                                               \
  int slot_ = (gridDim.x - 1 - blockIdx.x) * blockDim.x + threadIdx.x;                \
  bool live_ = slot_ < (k).n;                                                         \
  int g = env * (k).n + (live_ ? slot_ : 0);                                          \
  int gp = env * (k).n_pad + (slot_ < (k).n_pad ? slot_ : 0);

// Persistent particle kernels (round 2): the grid is one CTA per resident CTA slot and every WARP walks the 32-particle
// tiles w, w + W, w + 2W, ... (W = warps in the grid), prefetching its next tile while it computes the current one.
struct TileWalk {
  int lane, nwarps, tpe, ntiles, t;   // tpe = tiles per env; t = current tile (global over envs)
  int dq, dr;                         // nwarps = dq * tpe + dr: the walk's stride in (env, tile of env) coordinates
};
// a tile as (env, tile of the env): the walk advances it by additions (an integer division per tile and per lookahead
// was 4 % of k_p2g_pers' instructions and of its stall samples)
struct TilePos {
  int t, env, rem;
};
__device__ __forceinline__ TileWalk tile_walk(const MpmConst& k) {
  TileWalk w;
  w.lane = threadIdx.x & 31;
  w.nwarps = gridDim.x * (blockDim.x >> 5);
  w.tpe = k.n_pad >> 5;
  w.ntiles = k.B * w.tpe;
  w.t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  w.dq = w.nwarps / w.tpe;
  w.dr = w.nwarps - w.dq * w.tpe;
  return w;
}
__device__ __forceinline__ TilePos tile_first(const TileWalk& w) {
  TilePos p;
  p.t = w.t;
  p.env = w.t / w.tpe;
  p.rem = w.t - p.env * w.tpe;
  return p;
}
__device__ __forceinline__ TilePos tile_next(const TileWalk& w, TilePos p) {
  p.t += w.nwarps;
  p.env += w.dq;
  p.rem += w.dr;
  if (p.rem >= w.tpe) {
    p.rem -= w.tpe;
    ++p.env;
  }
  return p;
}
// particle of this lane in tile p: env, g = index into per-particle arrays, gp = index into the sorted tile arrays
__device__ __forceinline__ void tile_locate(const MpmConst& k, const TileWalk& w, const TilePos& p, int& env, int& g, int& gp,
                                            bool& live) {
  env = p.env;
  const int slot = p.rem * 32 + w.lane;
  live = slot < k.n;
  g = env * k.n + (live ? slot : 0);
  gp = env * k.n_pad + slot;
}
static inline int persistent_ctas(const MpmConst& k, int warps_per_cta, int ctas_per_sm) {
  return std::min(num_sms() * ctas_per_sm, cdiv((long long)k.B * (k.n_pad >> 5), warps_per_cta));
}
// fire-and-forget prefetch into L2 of this lane's 16 bytes of NQ quads of tile array `base` (a warp covers whole lines)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int Q0, int Q1, int NQ>
__device__ __forceinline__ void prefetch_quads(const float* __restrict__ base, int gp) {
  const float4* t = reinterpret_cast<const float4*>(base) + quad_index(gp, 0, NQ);
#pragma unroll
  for (int q = Q0; q <= Q1; ++q) prefetch_l2(t + q * 32);
}

// ------------------------------------------------------------------------------------------------
// Binning: key = 4x4x4-block-major cell key of base = int32(x*inv_dx - 0.5) (mpm_simulator.py:233),
// per-env stable counting sort.  Counts/starts are integer-atomic (deterministic); slots inside a cell
// are first claimed with atomics and then ranked by original index, which makes the permutation the
// stable argsort of the key, independent of scheduling.
// ------------------------------------------------------------------------------------------------
UD_DEV int cell_key(const MpmConst& k, const int base[3]) {
  int bx = min(max(base[0], 0), k.rx - 1), by = min(max(base[1], 0), k.ry - 1), bz = min(max(base[2], 0), k.rz - 1);
  int blk = ((bx >> 2) * k.nby + (by >> 2)) * k.nbz + (bz >> 2);
  return (blk << 6) | ((bx & 3) << 4) | ((by & 3) << 2) | (bz & 3);
}

__global__ void k_keys(MpmConst k, const float* __restrict__ x_aos, int32_t* __restrict__ keys,
                       int32_t* __restrict__ count, int32_t* __restrict__ out_base) {
  UD_PARTICLE_INDEX(k, env, g);
  if (!live_) return;
  float x[3] = {nan_to_num(x_aos[3 * g]), nan_to_num(x_aos[3 * g + 1]), nan_to_num(x_aos[3 * g + 2])};
  Stencil st;
  make_stencil(x, k.inv_dx, st);
  int key = cell_key(k, st.base);
  keys[g] = key;
  if (out_base) {
    out_base[3 * g] = st.base[0];
    out_base[3 * g + 1] = st.base[1];
    out_base[3 * g + 2] = st.base[2];
  }
  atomicAdd(&count[(size_t)env * (k.NK + 1) + key], 1);
}

// Exclusive scan of the NK per-env key counts in place (count -> start), start[NK] = n.  Two levels, every CTA owns a
// 1024-key chunk of one env: k_scan_sums leaves each chunk's total in `chunk_sum`, k_scan_chunks adds the totals of
// the chunks before its own to a local scan.  (Round 1 ran ONE 1024-thread CTA per env serially over its NK = 73 728
// keys: 74 us on 32 of 148 SMs.)
constexpr int SCAN_CHUNK = 1024;
__device__ __forceinline__ int block_sum_1024(int v, int* sh) {   // blockDim.x == 1024; returns the total to every thread
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  int t = sh[lane];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
  return t;
}
__global__ void __launch_bounds__(SCAN_CHUNK)
k_scan_sums(MpmConst k, const int32_t* __restrict__ cell_start, int32_t* __restrict__ chunk_sum) {
  __shared__ int sh[32];
  const int env = blockIdx.y, i = blockIdx.x * SCAN_CHUNK + threadIdx.x;
  const int v = i < k.NK ? cell_start[(size_t)env * (k.NK + 1) + i] : 0;
  const int t = block_sum_1024(v, sh);
  if (threadIdx.x == 0) chunk_sum[env * gridDim.x + blockIdx.x] = t;
}
__global__ void __launch_bounds__(SCAN_CHUNK)
k_scan_chunks(MpmConst k, int32_t* __restrict__ cell_start, const int32_t* __restrict__ chunk_sum) {
  __shared__ int sh[32], wsum[32];
  const int env = blockIdx.y, i = blockIdx.x * SCAN_CHUNK + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int32_t* c = cell_start + (size_t)env * (k.NK + 1);
  // totals of the chunks before mine (at most a few hundred values, strided over the block)
  int before = 0;
  for (int j = threadIdx.x; j < (int)blockIdx.x; j += SCAN_CHUNK) before += chunk_sum[env * gridDim.x + j];
  before = block_sum_1024(before, sh);
  const int v = i < k.NK ? c[i] : 0;
  int incl = v;   // inclusive scan inside the warp, then across the 32 warps
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, off);
    incl += lane >= off ? t : 0;
  }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  int wt = wsum[lane];
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, wt, off);
    wt += lane >= off ? t : 0;
  }
  const int warp_before = __shfl_sync(0xffffffffu, wt, wid) - wsum[wid];
  if (i < k.NK) c[i] = before + warp_before + incl - v;
  if (i == k.NK - 1) c[k.NK] = before + warp_before + incl;
}

__global__ void k_place(MpmConst k, const int32_t* __restrict__ keys, const int32_t* __restrict__ cell_start,
                        int32_t* __restrict__ cursor, int32_t* __restrict__ tmp_idx) {
  UD_PARTICLE_INDEX(k, env, g);
  if (!live_) return;
  int key = keys[g];
  int slot = cell_start[(size_t)env * (k.NK + 1) + key] + atomicAdd(&cursor[(size_t)env * k.NK + key], 1);
  tmp_idx[(size_t)env * k.n + slot] = g - env * k.n;
}

constexpr int RANK_BIG = 160;          // occupancy above which a cell is ranked by k_rank_big
constexpr int RANK_BIG_BLOCK = 256;
constexpr int RANK_BIG_WORDS = 4096;   // bitmap words per CTA: index spread up to 131 072 inside one cell
__global__ void k_rank(MpmConst k, const int32_t* __restrict__ keys, const int32_t* __restrict__ cell_start,
                       const int32_t* __restrict__ tmp_idx, int32_t* __restrict__ perm, int32_t* __restrict__ inv_perm,
                       int32_t* __restrict__ big_list, int32_t* __restrict__ big_count) {
  UD_PARTICLE_INDEX(k, env, g);
  if (!live_) return;
  int p = tmp_idx[g];
  int key = keys[env * k.n + p];
  const int32_t* cs = cell_start + (size_t)env * (k.NK + 1);
  int lo = cs[key], hi = cs[key + 1];
  const int32_t* seg = tmp_idx + (size_t)env * k.n;
  if (hi - lo > RANK_BIG) {   // a crowded cell: the count below is quadratic in the occupancy -> k_rank_big
    if (g - env * k.n == lo) {
      const int e = atomicAdd(big_count, 1);
      big_list[2 * e] = env;
      big_list[2 * e + 1] = key;
    }
    return;
  }
  int rank = 0;
  for (int j = lo; j < hi; ++j) rank += seg[j] < p;
  perm[(size_t)env * k.n + lo + rank] = p;
  inv_perm[(size_t)env * k.n + p] = lo + rank;   // sorted slot of original particle p (the un-sort gathers by it)
}

// Stable ranks inside the crowded cells k_rank listed (a rope or a settled pile puts thousands of particles into one
// cell: whip_rope at add_box density 25 has 2 025 per cell, where the per-particle count took 1.2 ms per sort).  One CTA
// per listed cell: the cell's particle indices set bits in a shared-memory bitmap over [min index, max index], a
// popcount scan of the bitmap words gives every index its rank among the cell's indices: O(occupancy + index spread / 32).
// A cell whose indices spread over more than 32 * RANK_BIG_WORDS falls back to the counting loop, spread over the CTA.
__global__ void __launch_bounds__(RANK_BIG_BLOCK)
k_rank_big(MpmConst k, const int32_t* __restrict__ cell_start, const int32_t* __restrict__ tmp_idx,
           const int32_t* __restrict__ big_list, const int32_t* __restrict__ big_count, int32_t* __restrict__ perm,
           int32_t* __restrict__ inv_perm) {
  __shared__ unsigned bits[RANK_BIG_WORDS];
  __shared__ int before[RANK_BIG_WORDS];
  __shared__ int red[2][RANK_BIG_BLOCK / 32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nbig = *big_count;
  for (int e = blockIdx.x; e < nbig; e += gridDim.x) {
    const int env = big_list[2 * e], key = big_list[2 * e + 1];
    const int32_t* cs = cell_start + (size_t)env * (k.NK + 1);
    const int lo = cs[key], hi = cs[key + 1];
    const int32_t* seg = tmp_idx + (size_t)env * k.n;
    int32_t* pm = perm + (size_t)env * k.n;
    int32_t* ip = inv_perm + (size_t)env * k.n;
    int mn = INT_MAX, mx = -1;
    for (int j = lo + tid; j < hi; j += RANK_BIG_BLOCK) {
      const int p = seg[j];
      mn = min(mn, p);
      mx = max(mx, p);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    __syncthreads();   // the previous cell's readers of bits / before / red are done
    if (lane == 0) {
      red[0][wid] = mn;
      red[1][wid] = mx;
    }
    __syncthreads();
    mn = red[0][0];
    mx = red[1][0];
#pragma unroll
    for (int i = 1; i < RANK_BIG_BLOCK / 32; ++i) {
      mn = min(mn, red[0][i]);
      mx = max(mx, red[1][i]);
    }
    const int w0 = mn >> 5, nw = (mx >> 5) - w0 + 1;
    if (nw > RANK_BIG_WORDS) {   // CTA-uniform
      for (int j = lo + tid; j < hi; j += RANK_BIG_BLOCK) {
        const int p = seg[j];
        int rank = 0;
        for (int i = lo; i < hi; ++i) rank += seg[i] < p;
        pm[lo + rank] = p;
        ip[p] = lo + rank;
      }
      continue;
    }
    for (int i = tid; i < nw; i += RANK_BIG_BLOCK) bits[i] = 0u;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int j = lo + tid; j < hi; j += RANK_BIG_BLOCK) {
      const int p = seg[j];
      atomicOr(&bits[(p >> 5) - w0], 1u << (p & 31));
    }
    __syncthreads();
    for (int base = 0; base < nw; base += RANK_BIG_BLOCK) {   // exclusive popcount scan, RANK_BIG_BLOCK words per round
      const int i = base + tid;
      const int v = i < nw ? __popc(bits[i]) : 0;
      int incl = v;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        incl += lane >= off ? t : 0;
      }
      if (lane == 31) red[0][wid] = incl;
      __syncthreads();
      int wb = 0, tot = 0;
#pragma unroll
      for (int q = 0; q < RANK_BIG_BLOCK / 32; ++q) {
        const int t = red[0][q];
        wb += q < wid ? t : 0;
        tot += t;
      }
      const int carry = carry_s;
      if (i < nw) before[i] = carry + wb + incl - v;
      __syncthreads();
      if (tid == 0) carry_s = carry + tot;
      __syncthreads();
    }
    for (int j = lo + tid; j < hi; j += RANK_BIG_BLOCK) {
      const int p = seg[j];
      const int wi = (p >> 5) - w0;
      const int rank = before[wi] + __popc(bits[wi] & ((1u << (p & 31)) - 1u));
      pm[lo + rank] = p;
      ip[p] = lo + rank;
    }
  }
}

__global__ void k_identity_perm(MpmConst k, int32_t* __restrict__ perm, int32_t* __restrict__ inv_perm) {
  UD_PARTICLE_INDEX(k, env, g);
  if (live_) perm[g] = inv_perm[g] = g - env * k.n;
}

void launch_sort(const MpmConst& k, const float* x_aos, const MpmWs& ws, int32_t* out_base, cudaStream_t st) {
  if (!tuning_sort() && !out_base) {  // A/B: no binning
    KScope ks_(KC_SORT, st, 1);
    k_identity_perm<<<pgrid(k, 256), 256, 0, st>>>(k, ws.perm, ws.inv_perm);
    return;
  }
  KScope ks_(KC_SORT, st, 8);
  cudaMemsetAsync(ws.cell_start, 0, sizeof(int32_t) * (size_t)k.B * (k.NK + 1), st);
  cudaMemsetAsync(ws.cursor, 0, sizeof(int32_t) * ((size_t)k.B * k.NK + 1), st);   // + the crowded-cell counter
  k_keys<<<pgrid(k, 256), 256, 0, st>>>(k, x_aos, ws.keys, ws.cell_start, out_base);
  const dim3 sg(cdiv(k.NK, SCAN_CHUNK), k.B);
  k_scan_sums<<<sg, SCAN_CHUNK, 0, st>>>(k, ws.cell_start, ws.chunk_sum);
  k_scan_chunks<<<sg, SCAN_CHUNK, 0, st>>>(k, ws.cell_start, ws.chunk_sum);
  k_place<<<pgrid(k, 256), 256, 0, st>>>(k, ws.keys, ws.cell_start, ws.cursor, ws.tmp_idx);
  int32_t* big_count = ws.cursor + (size_t)k.B * k.NK;
  k_rank<<<pgrid(k, 256), 256, 0, st>>>(k, ws.keys, ws.cell_start, ws.tmp_idx, ws.perm, ws.inv_perm, ws.big_list, big_count);
  k_rank_big<<<2 * num_sms(), RANK_BIG_BLOCK, 0, st>>>(k, ws.cell_start, ws.tmp_idx, ws.big_list, big_count, ws.perm,
                                                       ws.inv_perm);
}

// ------------------------------------------------------------------------------------------------
// AoS (reference leaf layout) -> sorted SoA, with norm_grad_state's forward nan_to_num
// (mpm_simulator.py:376-381).
// ------------------------------------------------------------------------------------------------
__global__ void k_gather_state(MpmConst k, const float* __restrict__ x, const float* __restrict__ v,
                               const float* __restrict__ C, const float* __restrict__ F,
                               const int32_t* __restrict__ material, const float* __restrict__ h,
                               const int32_t* __restrict__ perm, float* __restrict__ ps,
                               int32_t* __restrict__ mat_s, float* __restrict__ h_s) {
  UD_PARTICLE_INDEX(k, env, g);
  if (slot_ >= k.n_pad) return;
  float st[PS_NCOMP];
  if (live_) {
    int p = perm[g];
    size_t o = (size_t)env * k.n + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) st[PS_X + c] = nan_to_num(x[3 * o + c]);
#pragma unroll
    for (int c = 0; c < 3; ++c) st[PS_V + c] = nan_to_num(v[3 * o + c]);
#pragma unroll
    for (int c = 0; c < 9; ++c) st[PS_C + c] = nan_to_num(C[9 * o + c]);
#pragma unroll
    for (int c = 0; c < 9; ++c) st[PS_F + c] = nan_to_num(F[9 * o + c]);
    mat_s[g] = material[p];
    h_s[g] = h[p];
  } else {  // padding of the env's last tile: a harmless particle (F = I) that dead lanes may read
#pragma unroll
    for (int c = 0; c < PS_NCOMP; ++c) st[c] = (c == PS_F || c == PS_F + 4 || c == PS_F + 8) ? 1.f : 0.f;
  }
  store_comps<0, PS_NCOMP, PS_NQ>(ps, gp, st);
}

void launch_gather_state(const MpmConst& k, const ud_mpm_state* in, const int32_t* material, const float* h,
                         const MpmWs& ws, float* ps_slot, cudaStream_t st) {
  KScope ks_(KC_GATHER, st);
  k_gather_state<<<pgrid(k, 256), 256, 0, st>>>(k, in->x, in->v, in->C, in->F, material, h, ws.perm, ps_slot,
                                                  ws.mat_s, ws.h_s);
}

// ------------------------------------------------------------------------------------------------
// Shared-memory staged scatter.  A CTA owns UD_BLOCK consecutive particles of the per-frame sort.
// Phase 0 (stage_group): rows are GROUPED by base cell inside the CTA.  At the start of a step the sort
// makes each cell one contiguous run (~3.4 runs per CTA in the plasticine scene), but the sort is per
// frame and particles cross cell faces during the 16 substeps: measured, the contiguous runs fragment to
// 10-16 per CTA by the end of a step while the DISTINCT cells stay at 4-6.  Grouping restores one
// segment per distinct cell, so the number of global REDs per CTA does not grow with fragmentation.
// Phase 1: every particle stages its node values at its grouped row (value-major, +1 padded columns).
// Phase 2 (stage_flush): thread (node, comp) sums its column over each segment in fixed row order and
// issues ONE 16-byte vector RED per (segment, node).
// ------------------------------------------------------------------------------------------------
template <int SB> constexpr int stg_pad() { return SB + 4; }  // multiple of 4: 16-byte aligned columns for LDS.128; +4 keeps 8 neighbouring
                                       // columns in distinct 16-byte bank groups
constexpr int RUN_TAB = 23;            // segments whose 27 target cells are tabulated (aliases the grouping scratch)
template <int SB>
struct StageMeta {
  int key[SB];            // by grouped row
  int base[SB][3];        // by grouped row
  int run_start[SB + 1];  // segment starts (grouped rows)
  int n_runs;
  int warp_cnt[SB / 32];
  union {
    struct {  // scratch of stage_group: one entry per (warp, distinct key in that warp)
      int ekey[SB];
      unsigned emask[SB];
      int efirst[SB];
      int egid[SB];
      int estart[SB];
    };
    int cell[RUN_TAB][27];  // after grouping: linear cell index of node j of segment r, or -1 (dropped)
  };
};
constexpr int DEAD_KEY = 0x40000000;
// Marks the 4x4x4 grid block of cell (ix,iy,iz) as scattered-into: a fire-and-forget store (a load-then-store would
// put an L2 round trip on every CTA's path to its first barrier: measured +20 us on k_p2g).
UD_DEV void mark_block(const MpmConst& k, int32_t* __restrict__ flag_env, int ix, int iy, int iz) {
  flag_env[((ix >> 2) * k.nby + (iy >> 2)) * k.nbz + (iz >> 2)] = BLK_MARK_CTA;
}
UD_DEV int base_key(const int base[3]) { return (base[0] * 2048 + base[1]) * 2048 + base[2]; }

// Groups the CTA's rows by key.  Returns this thread's grouped row; fills m.key / m.base / m.run_start /
// m.n_runs.  Deterministic: groups are ordered by first occurrence (warp, lane), rows inside a group by
// (warp, lane).  All threads of the CTA must call it; it ends with a barrier.
// `blk_flag` (nullable, already offset to the env): P2G marks the 4x4x4 grid blocks it scatters into, so that the grid
// update visits only those (k_grid_fwd).
template <bool CLAMP, int SB>
__device__ __forceinline__ int stage_group(const MpmConst& k, StageMeta<SB>& m, int key, const int base[3],
                                           int32_t* __restrict__ blk_flag = nullptr, int* my_run = nullptr) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned mm = __match_any_sync(0xffffffffu, key);   // lanes of my warp with my key
  const int leadlane = __ffs(mm) - 1;
  const bool lead = lane == leadlane;
  const unsigned lb = __ballot_sync(0xffffffffu, lead);
  if (lane == 0) m.warp_cnt[wid] = __popc(lb);
  __syncthreads();
  int ebase = 0, E = 0;
#pragma unroll
  for (int w = 0; w < SB / 32; ++w) {
    const int c = m.warp_cnt[w];
    ebase += w < wid ? c : 0;
    E += c;
  }
  const int e = ebase + __popc(lb & lt);   // entry index (meaningful on leader lanes)
  if (lead) {
    m.ekey[e] = key;
    m.emask[e] = mm;
  }
  __syncthreads();
  if (t < E) {  // first entry with the same key
    const int kk = m.ekey[t];
    int first = t;
    for (int j = t - 1; j >= 0; --j) first = m.ekey[j] == kk ? j : first;
    m.efirst[t] = first;
  }
  __syncthreads();
  if (t < E) {  // group id = number of group heads before my head
    const int first = m.efirst[t];
    int gid = 0;
    for (int j = 0; j < first; ++j) gid += m.efirst[j] == j;
    m.egid[t] = gid;
  }
  __syncthreads();
  if (t < E) {  // start row of my entry = rows of earlier groups + rows of earlier entries of my group
    const int gid = m.egid[t];
    int start = 0, ng = 0;
    for (int j = 0; j < E; ++j) {
      const int gj = m.egid[j];
      const int cnt = __popc(m.emask[j]);
      start += (gj < gid || (gj == gid && j < t)) ? cnt : 0;
      ng = max(ng, gj + 1);
    }
    m.estart[t] = start;
    if (m.efirst[t] == t) m.run_start[gid] = start;
    if (t == 0) {
      m.n_runs = ng;
      m.run_start[ng] = SB;
    }
  }
  __syncthreads();
  const int my_entry = __shfl_sync(0xffffffffu, e, leadlane);
  const int pos = m.estart[my_entry] + __popc(mm & lt);
  if (my_run) *my_run = m.egid[my_entry];   // segment of this thread's row (read before the scratch dies)
  m.key[pos] = key;
  m.base[pos][0] = base[0];
  m.base[pos][1] = base[1];
  m.base[pos][2] = base[2];
  __syncthreads();  // the scratch is dead from here on: its storage becomes the cell table
  const int nr = m.n_runs;
  const bool tab = nr <= RUN_TAB;
  if (tab || (!CLAMP && blk_flag)) {  // target cells of every (segment, node), computed once by all threads of the CTA
    for (int e = t; e < nr * 27; e += SB) {
      const int r = e / 27, j = e - r * 27;
      const int s0 = m.run_start[r];
      const int a = j / 9, b = (j / 3) % 3, c = j % 3;
      int ix, iy, iz;
      if (CLAMP) {
        ix = idx_gather(m.base[s0][0] + a, k.rx);
        iy = idx_gather(m.base[s0][1] + b, k.ry);
        iz = idx_gather(m.base[s0][2] + c, k.rz);
      } else {
        ix = idx_scatter(m.base[s0][0] + a, k.rx);
        iy = idx_scatter(m.base[s0][1] + b, k.ry);
        iz = idx_scatter(m.base[s0][2] + c, k.rz);
      }
      const bool dropped = (ix | iy | iz) < 0 || (m.key[s0] & DEAD_KEY);
      if (tab) m.cell[r][j] = dropped ? -1 : (ix * k.ry + iy) * k.rz + iz;
      // a 3-node span touches at most two blocks per axis, both reached by its end nodes: corners mark everything
      const bool corner = a != 1 && b != 1 && c != 1;
      if (!CLAMP && blk_flag && !dropped && (k.mark == 1 || (k.mark == 2 && corner))) mark_block(k, blk_flag, ix, iy, iz);
    }
  }
  return pos;  // callers barrier between staging and flush, which also publishes the table
}

// phase 2: NC (3 or 4) components per node, staged value-major as column (node_local*NC + comp).
// Thread t = node_local*4 + comp sums its column over each run; the 4 lanes of a node then assemble a
// float4 with three shuffles and lane comp==0 issues ONE 16-byte vector RED (REDG.E.ADD.F32x4) per
// (run, node): scalar REDs cost ~4x more L2 atomic work (measured: 76 of 236 us in k_p2g).
// The 27 nodes are staged in windows of NPH nodes ([j0, j0+nj)), which divides the shared-memory
// footprint (and multiplies the resident CTAs per SM) without changing the amount of flush work.
template <int NC, bool CLAMP, bool DET, int SB>
__device__ __forceinline__ void stage_flush(const MpmConst& k, const float* __restrict__ sv, const StageMeta<SB>& m,
                                            float4* __restrict__ genv, int j0, int nj) {
  const int t = threadIdx.x;
  if (t >= ((nj * 4 + 31) & ~31)) return;  // whole extra warps leave; partial warps stay for the shuffles
  const int jl = t >> 2, c = t & 3;
  const bool sums = jl < nj && c < NC;
  const int j = j0 + jl;
  const int a = j / 9, b = (j / 3) % 3, cc = j % 3;
  const float* col = sv + ((jl < nj ? jl : 0) * NC + (c < NC ? c : 0)) * stg_pad<SB>();
  const int nr = m.n_runs;
  const bool tab = nr <= RUN_TAB;  // block-uniform
  for (int r = 0; r < nr; ++r) {
    const int s0 = m.run_start[r], s1 = m.run_start[r + 1];
    if (m.key[s0] & DEAD_KEY) continue;  // block-uniform
    float acc = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;  // independent partial sums (fixed order)
    if (sums) {
      const float* pc = col + s0;
      int left = s1 - s0;
      int head = (4 - (s0 & 3)) & 3;  // rows up to a 16-byte boundary
      head = head < left ? head : left;
      left -= head;
#pragma unroll 1
      for (; head > 0; --head) acc += *pc++;
      const float4* p4 = reinterpret_cast<const float4*>(pc);
#pragma unroll 1
      for (; left >= 8; left -= 8) {  // body: two LDS.128 per 8 rows
        const float4 q0 = p4[0], q1 = p4[1];
        p4 += 2;
        acc += q0.x;
        acc1 += q0.y;
        acc2 += q0.z;
        acc3 += q0.w;
        acc += q1.x;
        acc1 += q1.y;
        acc2 += q1.z;
        acc3 += q1.w;
      }
      if (left >= 4) {
        const float4 q0 = *p4++;
        acc += q0.x;
        acc1 += q0.y;
        acc2 += q0.z;
        acc3 += q0.w;
        left -= 4;
      }
      pc = reinterpret_cast<const float*>(p4);
#pragma unroll 1
      for (; left > 0; --left) acc += *pc++;
      acc = (acc + acc1) + (acc2 + acc3);
    }
    float4 val;
    val.x = acc;
    val.y = __shfl_down_sync(0xffffffffu, acc, 1);
    val.z = __shfl_down_sync(0xffffffffu, acc, 2);
    val.w = __shfl_down_sync(0xffffffffu, acc, 3);
    if (c != 0 || jl >= nj) continue;
    int cell;
    if (tab) {
      cell = m.cell[r][j];
      if (cell < 0) continue;
    } else {
      int ix, iy, iz;
      if (CLAMP) {
        ix = idx_gather(m.base[s0][0] + a, k.rx);
        iy = idx_gather(m.base[s0][1] + b, k.ry);
        iz = idx_gather(m.base[s0][2] + cc, k.rz);
      } else {
        ix = idx_scatter(m.base[s0][0] + a, k.rx);
        iy = idx_scatter(m.base[s0][1] + b, k.ry);
        iz = idx_scatter(m.base[s0][2] + cc, k.rz);
        if ((ix | iy | iz) < 0) continue;
      }
      cell = (ix * k.ry + iy) * k.rz + iz;
    }
    if (DET) {  // UD_P2G_DETERMINISTIC: 4 integer REDs on 64-bit fixed point (associative => order-independent)
      unsigned long long* acc64 = reinterpret_cast<unsigned long long*>(genv) + 4 * (size_t)cell;
      atomicAdd(acc64 + 0, (unsigned long long)__double2ll_rn((double)val.x * FIX_SCALE));
      atomicAdd(acc64 + 1, (unsigned long long)__double2ll_rn((double)val.y * FIX_SCALE));
      atomicAdd(acc64 + 2, (unsigned long long)__double2ll_rn((double)val.z * FIX_SCALE));
      atomicAdd(acc64 + 3, (unsigned long long)__double2ll_rn((double)val.w * FIX_SCALE));
    } else {
      atomicAdd(&genv[cell], val);
    }
  }
}
template <int SB>
constexpr size_t stage_smem_bytes(int nc, int nph) { return sizeof(float) * nph * nc * stg_pad<SB>() + sizeof(StageMeta<SB>); }
constexpr int P2G_BLOCK = 64;    // threads (= particles) per CTA of k_p2g: with 14-node windows the 56 flush lanes are
                                 // the whole CTA, so no warp idles at a barrier while others sum columns
constexpr int G2PB_BLOCK = UD_BLOCK;  // k_g2p_bwd: one window of 27 nodes, 108 flush lanes of 128 (two windows cost it registers)
constexpr int P2G_NPH = 14;      // nodes per staging window of k_p2g (27 nodes -> 2 windows)
constexpr int G2PB_NPH = 27;
// float4 slots between the cells of a node tile: 33 (528 B), not 32, so that lanes in DIFFERENT cells reading the same
// node hit different banks (32 puts node j of every cell in the same four banks: an nseg-way conflict on each of the 27
// stencil loads; measured k_g2p 50.4 -> 46.0 us).  Only slots 0..26 of a cell are used, so 4 cells still fit in 4 * 32.
#ifndef UD_NT_STRIDE
#define UD_NT_STRIDE 33
#endif
constexpr int NT_STRIDE = UD_NT_STRIDE;
constexpr int G2P_TILE_CELLS = 4;
// One-tile-per-warp kernels: every warp asks L2 for the quads a warp one wave of CTAs later will load first (the CTAs
// are dispatched in tile order), so that warp's first dependent load is an L2 hit instead of a DRAM round trip.
// Measured: k_g2p_bwd_warp 113.5 -> 110.5 us; nothing for k_g2p (45.8 / 45.9), where it is not used.
#ifndef UD_AHEAD_TILES
#define UD_AHEAD_TILES (148 * 32)
#endif
template <int Q0, int Q1, int NQ>
__device__ __forceinline__ void ahead_prefetch(const MpmConst& k, const float* __restrict__ base, int gp) {
  if (UD_AHEAD_TILES > 0) {
    const int gpa = gp + UD_AHEAD_TILES * 32;
    if (gpa < k.N_pad) prefetch_quads<Q0, Q1, NQ>(base, gpa);
  }
}
static_assert((G2P_TILE_CELLS - 1) * NT_STRIDE + 27 <= G2P_TILE_CELLS * 32, "node tile: the padded cells must fit");    // distinct base cells per warp whose stencils k_g2p keeps in shared memory
constexpr int G2PB_TILE_RUNS = 12;  // segments whose 27 grid velocities k_g2p_bwd keeps in shared memory (5.2 KB: the CTA
                                    // stays at 4 per SM); CTAs with more distinct cells gather from L1/L2 as before
constexpr size_t g2pb_tile_offset() { return (stage_smem_bytes<G2PB_BLOCK>(3, G2PB_NPH) + 15) & ~(size_t)15; }
constexpr size_t g2pb_smem_bytes() { return g2pb_tile_offset() + sizeof(float4) * G2PB_TILE_RUNS * 27; }

// ------------------------------------------------------------------------------------------------
// P2G: F update + SVD + plasticity + stress (mpm_simulator.py:238-268), then the 27-node scatter of
// {momentum, mass} (p2g_micro, :178-194) as ONE 16-byte vector reduction per node
// (red.global.add.v4.f32 -> REDG.E.ADD.F32x4).  Out-of-range nodes are dropped (JAX scatter rule).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_particle(const float* __restrict__ ps, int gp, float x[3], float v[3], Mat3& C, Mat3& F) {
  float st[PS_NCOMP];
  load_comps<0, PS_NCOMP, PS_NQ>(ps, gp, st);   // 6 x LDG.128, tile base + immediate
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = st[PS_X + c];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = st[PS_V + c];
#pragma unroll
  for (int c = 0; c < 9; ++c) C.m[c] = st[PS_C + c];
#pragma unroll
  for (int c = 0; c < 9; ++c) F.m[c] = st[PS_F + c];
}

// Per-particle front half of P2G, shared by the scatter variants: load, F update, SVD (warm-started from the previous
// substep's V^T), plastic clip, stress; stores F' (+ V^T for the next warm start, or the whole SVD in the recompute
// pass).  Returns the stencil and the affine scatter value at node (a,b,c):
//   wt * (u + a*Ac[0] + b*Ac[1] + c*Ac[2]),  u = p_mass v - dx A fx,  Ac[j] = dx * column j of A
// `vt_in` is the forward's warm-start buffer (VT tile) or, in the recompute pass (svd_out != null), the previous
// substep's SVD tile.
// Every global load of the particle is issued before the first one is consumed (one DRAM round trip per warp, not
// three: ncu showed the warm-start and per-particle parameter loads waiting behind the stencil of x).
// everything P2G reads of one particle, as loaded (the persistent kernel holds the NEXT tile's loads in flight here)
struct P2gIn {
  float4 q[PS_NQ];   // state tile quads: x v C F
  float4 vt[3];      // warm start: V^T (9) of the previous substep
  float h, mu, la;
  int mat;
};
__device__ __forceinline__ void p2g_issue_loads(int env, int g, int gp, const float* __restrict__ ps_in,
                                                const float* __restrict__ mu_s, const float* __restrict__ la_s,
                                                const int32_t* __restrict__ mat_s, const float* __restrict__ h_s,
                                                const float* __restrict__ vt_in, bool vt_is_svd_tile, P2gIn& in) {
  const float4* t = reinterpret_cast<const float4*>(ps_in) + quad_index(gp, 0, PS_NQ);
#pragma unroll
  for (int q = 0; q < PS_NQ; ++q) in.q[q] = t[q * 32];
  if (vt_in) {   // grid-uniform
    if (vt_is_svd_tile) {   // V^T = components 12..20 of the SVD tile: quads 3, 4, 5
      const float4* v = reinterpret_cast<const float4*>(vt_in) + quad_index(gp, 0, SV_NQ);
#pragma unroll
      for (int q = 0; q < 3; ++q) in.vt[q] = v[(3 + q) * 32];
    } else {
      const float4* v = reinterpret_cast<const float4*>(vt_in) + quad_index(gp, 0, VT_NQ);
#pragma unroll
      for (int q = 0; q < 3; ++q) in.vt[q] = v[q * 32];
    }
  }
  in.h = h_s[g];
  in.mat = mat_s[g];
  in.mu = mu_s[env];
  in.la = la_s[env];
}
static_assert(SV_VT == 12 && PS_NQ == 6, "p2g_issue_loads reads V^T as whole quads");

// Per-particle front half of P2G, shared by the scatter variants: F update, SVD (warm-started from the previous
// substep's V^T), plastic clip, stress; stores F' (+ V^T for the next warm start, or the whole SVD in the recompute
// pass).  Returns the stencil and the affine scatter value at node (a,b,c):
//   wt * (u + a*Ac[0] + b*Ac[1] + c*Ac[2]),  u = p_mass v - dx A fx,  Ac[j] = dx * column j of A
template <bool LIQ>
__device__ __forceinline__ void p2g_front_loaded(const MpmConst& k, int gp, bool wr, bool live, const P2gIn& in, bool warm,
                                                 float* __restrict__ ps_out, float* __restrict__ vt_out,
                                                 float* __restrict__ svd_out, Stencil& st, float u[3], float Ac[3][3]) {
  const float sq[PS_NCOMP] = {in.q[0].x, in.q[0].y, in.q[0].z, in.q[0].w, in.q[1].x, in.q[1].y, in.q[1].z, in.q[1].w,
                              in.q[2].x, in.q[2].y, in.q[2].z, in.q[2].w, in.q[3].x, in.q[3].y, in.q[3].z, in.q[3].w,
                              in.q[4].x, in.q[4].y, in.q[4].z, in.q[4].w, in.q[5].x, in.q[5].y, in.q[5].z, in.q[5].w};
  float x[3], v[3];
  Mat3 C, F;
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = sq[PS_X + c];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = sq[PS_V + c];
#pragma unroll
  for (int c = 0; c < 9; ++c) C.m[c] = sq[PS_C + c];
#pragma unroll
  for (int c = 0; c < 9; ++c) F.m[c] = sq[PS_F + c];
  float vt0[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  if (warm) {
    // forward warm-start buffer: comps 0..8 of quads 0..2; SVD tile: comps 12..20 = quads 3..5 (same quad-relative layout)
    vt0[0] = in.vt[0].x; vt0[1] = in.vt[0].y; vt0[2] = in.vt[0].z; vt0[3] = in.vt[0].w;
    vt0[4] = in.vt[1].x; vt0[5] = in.vt[1].y; vt0[6] = in.vt[1].z; vt0[7] = in.vt[1].w;
    vt0[8] = in.vt[2].x;
  }
  make_stencil(x, k.inv_dx, st);
  Consti o;
  constitutive_pre(k, C, F, in.mu, in.la, in.h, in.mat, o);
  // A liquid particle needs no factorisation (constitutive_post_liquid).  A warp of liquid particles skips the SVD and
  // the V^T / SVD stores (every caller runs whole warps through here); in a mixed warp the liquid lanes ride along
  // and ignore the result.  P2G^T decides per particle the same way and never reads a liquid particle's SVD.
  const bool liq = LIQ && o.liquid;
  const bool no_svd = LIQ && __all_sync(0xffffffffu, liq || !live);
  if (!no_svd) svd3_ws(o.F1, o.U, o.s, o.Vt, warm, vt0);
#ifdef UD_NO_PLASTIC_FAST
  const bool all_plastic = false;
#else
  const bool all_plastic = __all_sync(0xffffffffu, o.plastic);
#endif
  // Only warp-uniform branches here: a per-lane branch on `liq` ended with nvcc 12.9 re-materialising the vote of
  // `all_plastic` inside the non-liquid side (VOTE.ALL behind WARPSYNC.COLLECTIVE in divergent code: a warp mixing liquid
  // with other particles deadlocked).  A mixed warp therefore runs the general path on every lane and the liquid lanes
  // then take their own J = |det F1| through selects, so that forward and P2G^T agree on it bit for bit.
  if (no_svd) {
    constitutive_post_liquid(k, C, o);
    vt_out = svd_out = nullptr;
  } else {
    constitutive_post(k, C, o, all_plastic);
    if (LIQ) {
      // the general path already gave a liquid lane F2 = F1 and affine = p_mass C off the diagonal (mu = 0); only the
      // isotropic term changes with J (constitutive_post_liquid's expression)
      const float Jd = fabsf(det3_pivoted(o.F1));
      const float iso_l = o.la * Jd * (Jd - 1.f);
      const float cs = k.c_stress_mul / k.c_stress_div;
#pragma unroll
      for (int i = 0; i < 3; ++i) o.affine(i, i) = liq ? cs * iso_l + k.p_mass * C(i, i) : o.affine(i, i);
    }
  }
  if (wr) {   // padding lanes of an env's last tile write too: the next substep's dead lanes read initialised memory
    store_comps<PS_F, 9, PS_NQ>(ps_out, gp, o.F2.m);
    if (vt_out) {
      float t[VT_NCOMP];
#pragma unroll
      for (int c = 0; c < VT_NCOMP; ++c) t[c] = c < 9 ? o.Vt.m[c] : 0.f;
      store_comps<0, VT_NCOMP, VT_NQ>(vt_out, gp, t);
    }
    if (svd_out) {  // recompute pass of the adjoint: keep the SVD so that P2G^T does not redo it
      float t[SV_NCOMP];
#pragma unroll
      for (int c = 0; c < 9; ++c) t[SV_U + c] = o.U.m[c];
#pragma unroll
      for (int c = 0; c < 3; ++c) t[SV_S + c] = o.s[c];
#pragma unroll
      for (int c = 0; c < 9; ++c) t[SV_VT + c] = o.Vt.m[c];
      t[21] = t[22] = t[23] = 0.f;
      store_comps<0, SV_NCOMP, SV_NQ>(svd_out, gp, t);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) Ac[j][i] = k.dx * o.affine(i, j);
    u[i] = k.p_mass * v[i] - (Ac[0][i] * st.fx[0] + Ac[1][i] * st.fx[1] + Ac[2][i] * st.fx[2]);
  }
}
// load + front half (the one-tile-per-warp kernels): every global load of the particle is issued before the first one
// is consumed (one DRAM round trip per warp, not three)
template <bool LIQ>
__device__ __forceinline__ void p2g_front(const MpmConst& k, int env, int g, int gp, bool wr, bool live, const float* __restrict__ ps_in,
                                          float* __restrict__ ps_out, const float* __restrict__ mu_s,
                                          const float* __restrict__ la_s, const int32_t* __restrict__ mat_s,
                                          const float* __restrict__ h_s, const float* __restrict__ vt_in,
                                          float* __restrict__ vt_out, float* __restrict__ svd_out, bool vt_svd,
                                          Stencil& st, float u[3], float Ac[3][3]) {
  P2gIn in;
  p2g_issue_loads(env, g, gp, ps_in, mu_s, la_s, mat_s, h_s, vt_in, vt_svd, in);
  p2g_front_loaded<LIQ>(k, gp, wr, live, in, vt_in != nullptr, ps_out, vt_out, svd_out, st, u, Ac);
}

// MODE 0: CTA-staged scatter + fp32 vector REDs; 1: staged + 64-bit fixed-point REDs (deterministic);
//      2: A/B baseline -- every particle issues its 27 vector REDs itself (no shared memory, no grouping)
// (round-1 kernel, kept as the A/B partner of k_p2g_warp: ud_tuning_set("warp", 0))
template <int MODE>
__global__ void __launch_bounds__(P2G_BLOCK, 12)   // 80 registers: 12 CTAs per SM, what the staging shared memory allows too
k_p2g(MpmConst k, const float* ps_in, float* ps_out, float4* __restrict__ grid,
      const float* __restrict__ mu_s, const float* __restrict__ la_s, const int32_t* __restrict__ mat_s,
      const float* __restrict__ h_s, const float* __restrict__ vt_in, float* __restrict__ vt_out,
      float* __restrict__ svd_out, int vt_svd, int32_t* __restrict__ blk_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sv = reinterpret_cast<float*>(smem_raw);                               // [P2G_NPH*4][STG_PAD]
  constexpr int STG_PAD = stg_pad<P2G_BLOCK>();
  StageMeta<P2G_BLOCK>& meta = *reinterpret_cast<StageMeta<P2G_BLOCK>*>(sv + P2G_NPH * 4 * STG_PAD);
  UD_PARTICLE_INDEX(k, env, g);
  Stencil st;
  float u[3], Ac[3][3];
  p2g_front<false>(k, env, g, gp, slot_ < k.n_pad, live_, ps_in, ps_out, mu_s, la_s, mat_s, h_s, vt_in, vt_out, svd_out, vt_svd != 0, st, u, Ac);
  constexpr bool DET = MODE == 1;
  int row = 0;
  if (MODE != 2)
    row = stage_group<false, P2G_BLOCK>(k, meta, live_ ? (base_key(st.base) & ~DEAD_KEY) : DEAD_KEY, st.base,
                                        blk_flag + (size_t)env * (k.nbx * k.nby * k.nbz));
  const float lw = live_ ? 1.f : 0.f;
  if constexpr (MODE == 2) {
    float4* genv = grid + (size_t)env * k.G;
    if (live_) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int ix = idx_scatter(st.base[0] + a, k.rx), iy = idx_scatter(st.base[1] + b, k.ry),
                      iz = idx_scatter(st.base[2] + c, k.rz);
            if ((ix | iy | iz) < 0) continue;
            mark_block(k, blk_flag + (size_t)env * (k.nbx * k.nby * k.nbz), ix, iy, iz);
            const float wt = st.w[a][0] * st.w[b][1] * st.w[c][2];
            float4 val;
            val.x = wt * (u[0] + (float)a * Ac[0][0] + (float)b * Ac[1][0] + (float)c * Ac[2][0]);
            val.y = wt * (u[1] + (float)a * Ac[0][1] + (float)b * Ac[1][1] + (float)c * Ac[2][1]);
            val.z = wt * (u[2] + (float)a * Ac[0][2] + (float)b * Ac[1][2] + (float)c * Ac[2][2]);
            val.w = wt * k.p_mass;
            atomicAdd(&genv[(ix * k.ry + iy) * k.rz + iz], val);
          }
    }
  } else {
#pragma unroll
  for (int j0 = 0; j0 < 27; j0 += P2G_NPH) {
    if (j0) __syncthreads();  // the previous window has been flushed
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float wa = st.w[a][0] * lw;
      float ua[3] = {u[0] + (float)a * Ac[0][0], u[1] + (float)a * Ac[0][1], u[2] + (float)a * Ac[0][2]};
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const float wab = wa * st.w[b][1];
        float uab[3] = {ua[0] + (float)b * Ac[1][0], ua[1] + (float)b * Ac[1][1], ua[2] + (float)b * Ac[1][2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int j = a * 9 + b * 3 + c;
          if (j < j0 || j >= j0 + P2G_NPH) continue;  // compile-time after unrolling
          const float wt = wab * st.w[c][2];
          float* dst = sv + ((j - j0) * 4) * STG_PAD + row;
          dst[0] = wt * (uab[0] + (float)c * Ac[2][0]);
          dst[STG_PAD] = wt * (uab[1] + (float)c * Ac[2][1]);
          dst[2 * STG_PAD] = wt * (uab[2] + (float)c * Ac[2][2]);
          dst[3 * STG_PAD] = wt * k.p_mass;
        }
      }
    }
    __syncthreads();
    // DET: `grid` is the int64 accumulator array (32 B per cell = 2 float4 slots per cell)
    stage_flush<4, false, DET, P2G_BLOCK>(k, sv, meta, grid + (size_t)env * k.G * (DET ? 2 : 1), j0,
                               (27 - j0) < P2G_NPH ? (27 - j0) : P2G_NPH);
  }
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-local staged scatter (round 2).  A WARP owns one tile of 32 consecutive particles of the per-frame sort and
// never talks to another warp: no block barrier, no cross-warp merge.
//   group : __match_any_sync on the base cell -> segments (lanes sharing a cell); rows of a segment are made contiguous
//           (segments in leader-lane order, rows in lane order), so the reduction order is fixed
//   stage : every lane writes its node values as float4 {px,py,pz,m} at tile[row][node] (one STS.128 per node; the
//           432-byte row stride is conflict-free for 16-byte accesses)
//   flush : lane = node sums its node's float4 column over the rows of each segment (LDS.128 + 4 FADD per row) and
//           issues ONE 16-byte vector RED per (segment, node)
// One window of all 27 nodes (13.8 KB per warp).  Two windows of 14 / 13 nodes and three of 9 (more resident warps, one
// flush pass per window) were measured and lost: 13.76 / 16.34 vs 12.79 ms per step (profiles/r02_summary.md).
// ------------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;
// The 4x4x4 grid blocks a substep's P2G scatters into, listed by the scatter itself: the first warp to touch a block
// stamps its flag (flags are zeroed once per call, the stamp is substep + 1) and appends it to the substep's list, which
// the grid update then walks.  The flag is read through L1 first: after the first touch per SM the check is an L1 hit
// and the atomic is skipped (a stale "unmarked" only costs a redundant atomic, never a duplicate entry).
struct BlkList {
  int32_t* flag;    // [B * nblk]
  int32_t* list;    // this substep's list
  int32_t* count;   // this substep's counter; count[1]: set when a particle's stencil leaves the grid (mpm_internal.h)
  int stamp;
};
struct WarpGroup {
  unsigned mm;   // lanes whose particle sits in my base cell
  unsigned lb;   // leader lanes, one per distinct cell of the warp
  int row;       // my staging row
  int nseg;      // distinct cells (segments) of the warp
};
// Per-warp segment table (shared memory, 32 x int2), written by the leader lanes while grouping and read by the flush
// with one broadcast LDS.64 per segment instead of five shuffles + index arithmetic:
//   .x = first row | rows << 6 | leader lane << 12 | live << 17 | interior << 18
//   .y = linear cell index of the segment's base node (valid when interior: all 27 nodes in range, no wrap / clamp)
constexpr int SEG_TABLE_BYTES = 32 * sizeof(int2);
__device__ __forceinline__ WarpGroup warp_group(const MpmConst& k, int key, const int base[3], bool live, int2* __restrict__ seg) {
  const int lane = threadIdx.x & 31;
  WarpGroup g;
  g.mm = __match_any_sync(FULL, key);
  const int lead = __ffs(g.mm) - 1;
  g.lb = __ballot_sync(FULL, lane == lead);
  g.nseg = __popc(g.lb);
  int start = 0;
  for (unsigned b = g.lb; b; b &= b - 1) {   // warp-uniform trip count = distinct cells (2-4 in the plasticine scene)
    const int L = __ffs(b) - 1;
    const int cnt = __popc(__shfl_sync(FULL, g.mm, L));
    start += L < lead ? cnt : 0;
  }
  g.row = start + __popc(g.mm & ((1u << lane) - 1u));
  __syncwarp();   // persistent kernels: every lane has finished reading the previous tile's table
  if (lane == lead) {
    const bool interior = base[0] >= 0 && base[0] + 2 < k.rx && base[1] >= 0 && base[1] + 2 < k.ry && base[2] >= 0 &&
                          base[2] + 2 < k.rz;
    seg[__popc(g.lb & ((1u << lane) - 1u))] =
        make_int2(start | (__popc(g.mm) << 6) | (lane << 12) | ((int)live << 17) | ((int)interior << 18),
                  (base[0] * k.ry + base[1]) * k.rz + base[2]);
  }
  return g;   // callers __syncwarp() between staging and flush, which also publishes the table
}
constexpr int WARP_TILE_NODES = 27;
constexpr size_t WARP_TILE_BYTES = sizeof(float4) * 32 * WARP_TILE_NODES + SEG_TABLE_BYTES;   // staging rows + segment table

// Flush: `tile` is this warp's [32][27] float4 array (432-byte rows: conflict-free 16-byte accesses).  Lane = node sums
// its column over the rows of each segment in row order (two accumulators, eight rows in flight) and issues ONE 16-byte
// vector RED per (segment, node).
template <bool CLAMP, bool DET>
__device__ __forceinline__ void warp_flush(const MpmConst& k, const float4* __restrict__ tile, const int2* __restrict__ seg,
                                           const WarpGroup& g, const int base[3], float4* __restrict__ genv,
                                           float wscale = 1.f) {
  constexpr int WS = WARP_TILE_NODES;
  const int lane = threadIdx.x & 31;
  const bool act = lane < 27;
  const int j = act ? lane : 0;
  const int a = j / 9, b = (j / 3) % 3, c = j % 3;
  const int lane_off = (a * k.ry + b) * k.rz + c;
  for (int s = 0; s < g.nseg; ++s) {
    const int2 e = seg[s];   // broadcast
    if (!((e.x >> 17) & 1)) continue;   // warp-uniform: the padding lanes of an env's last tile
    const int r_begin = e.x & 63, cnt = (e.x >> 6) & 63;
    int cell;
    bool ok = act;
    if ((e.x >> 18) & 1) {
      cell = e.y + lane_off;
    } else {   // a segment at the grid boundary: JAX's scatter (drop) / gather (clamp) index rules, node by node
      const int L = (e.x >> 12) & 31;
      const int bx = __shfl_sync(FULL, base[0], L), by = __shfl_sync(FULL, base[1], L), bz = __shfl_sync(FULL, base[2], L);
      int ix, iy, iz;
      if (CLAMP) {
        ix = idx_gather(bx + a, k.rx);
        iy = idx_gather(by + b, k.ry);
        iz = idx_gather(bz + c, k.rz);
      } else {
        ix = idx_scatter(bx + a, k.rx);
        iy = idx_scatter(by + b, k.ry);
        iz = idx_scatter(bz + c, k.rz);
      }
      ok = act && (ix | iy | iz) >= 0;
      cell = (ix * k.ry + iy) * k.rz + iz;
    }
    if (!ok) continue;
    // packed fp32 adds (FADD2, sm_100): the column sum is 2 issue slots per row instead of 4; same order, same
    // rounding as the scalar form
    float2 aL = make_float2(0.f, 0.f), aH = aL, bL = aL, bH = aL;
    const float4* p = tile + r_begin * WS + j;
    int left = cnt;
#define UD_ACC(L_, H_, q_) { L_ = __fadd2_rn(L_, make_float2(q_.x, q_.y)); H_ = __fadd2_rn(H_, make_float2(q_.z, q_.w)); }
    for (; left >= 8; left -= 8) {
      const float4 q0 = p[0], q1 = p[WS], q2 = p[2 * WS], q3 = p[3 * WS];
      const float4 q4 = p[4 * WS], q5 = p[5 * WS], q6 = p[6 * WS], q7 = p[7 * WS];
      p += 8 * WS;
      UD_ACC(aL, aH, q0) UD_ACC(bL, bH, q4) UD_ACC(aL, aH, q1) UD_ACC(bL, bH, q5)
      UD_ACC(aL, aH, q2) UD_ACC(bL, bH, q6) UD_ACC(aL, aH, q3) UD_ACC(bL, bH, q7)
    }
    if (left >= 4) {
      const float4 q0 = p[0], q1 = p[WS], q2 = p[2 * WS], q3 = p[3 * WS];
      p += 4 * WS;
      left -= 4;
      UD_ACC(aL, aH, q0) UD_ACC(bL, bH, q1) UD_ACC(aL, aH, q2) UD_ACC(bL, bH, q3)
    }
    for (; left > 0; --left) {
      const float4 q0 = *p;
      p += WS;
      UD_ACC(aL, aH, q0)
    }
#undef UD_ACC
    aL = __fadd2_rn(aL, bL);
    aH = __fadd2_rn(aH, bH);
    float4 acc = make_float4(aL.x, aL.y, aH.x, aH.y);
    acc.w *= wscale;   // P2G stages the bare weight in .w: p_mass is applied once per (segment, node), not per particle
    if (DET) {  // UD_P2G_DETERMINISTIC: 4 integer REDs on 64-bit fixed point (associative => order-independent)
      unsigned long long* acc64 = reinterpret_cast<unsigned long long*>(genv) + 4 * (size_t)cell;
      atomicAdd(acc64 + 0, (unsigned long long)__double2ll_rn((double)acc.x * FIX_SCALE));
      atomicAdd(acc64 + 1, (unsigned long long)__double2ll_rn((double)acc.y * FIX_SCALE));
      atomicAdd(acc64 + 2, (unsigned long long)__double2ll_rn((double)acc.z * FIX_SCALE));
      atomicAdd(acc64 + 3, (unsigned long long)__double2ll_rn((double)acc.w * FIX_SCALE));
    } else {
      atomicAdd(&genv[cell], acc);
    }
  }
}

// Marks (and lists) the 4x4x4 grid blocks the warp's scatter touches.  A 3-node span touches at most two blocks per
// axis, both reached by its end nodes, so the 8 corner nodes of every segment mark everything.  One lane per
// (segment, corner), four segments per pass, AFTER the flush: the flag loads of a warp are one round trip instead of
// one per segment inside the flush loop (ncu: 9 % of k_p2g_warp's stall samples sat on that dependent load).
// Measured: 147.6 -> 140.0 us per launch; issuing the flag loads before the staging instead was slower (156.7 us).
// The flag is read through L1 first: after the first touch per SM the check is an L1 hit and the atomic is skipped (a
// stale "unmarked" only costs a redundant atomic, never a duplicate entry).
__device__ __forceinline__ void warp_mark_blocks(const MpmConst& k, const WarpGroup& g, const int base[3], bool live,
                                                 const BlkList& bl, int env) {
  if (k.mark == 0) return;
  const int lane = threadIdx.x & 31;
  // a stencil that leaves the grid reads (G2P, clamped / wrapped indices) face cells of blocks nobody scattered into:
  // tell the grid update that its shell job is needed this substep (rare: a plain idempotent store)
  const bool outside = live && !(base[0] >= 0 && base[0] + 2 < k.rx && base[1] >= 0 && base[1] + 2 < k.ry && base[2] >= 0 &&
                                 base[2] + 2 < k.rz);
  if (__any_sync(FULL, outside) && lane == 0) bl.count[1] = 1;
  const int a = (lane & 4) ? 2 : 0, b = (lane & 2) ? 2 : 0, c = (lane & 1) ? 2 : 0;
  const int nseg = __popc(g.lb);
  for (int s0 = 0; s0 < nseg; s0 += 4) {   // warp-uniform; warps spanning more than four cells are rare
    const int my = s0 + (lane >> 3);
    int L = 0, i = 0;
    for (unsigned bits = g.lb; bits; bits &= bits - 1, ++i) L = i == my ? __ffs(bits) - 1 : L;
    const int bx = __shfl_sync(FULL, base[0], L), by = __shfl_sync(FULL, base[1], L), bz = __shfl_sync(FULL, base[2], L);
    const int lv = __shfl_sync(FULL, (int)live, L);
    if (my >= nseg || !lv) continue;
    const int ix = idx_scatter(bx + a, k.rx), iy = idx_scatter(by + b, k.ry), iz = idx_scatter(bz + c, k.rz);
    if ((ix | iy | iz) < 0) continue;
    const int w = env * (k.nbx * k.nby * k.nbz) + ((ix >> 2) * k.nby + (iy >> 2)) * k.nbz + (iz >> 2);
    if (__ldca(bl.flag + w) != bl.stamp && atomicExch(bl.flag + w, bl.stamp) != bl.stamp) bl.list[atomicAdd(bl.count, 1)] = w;
  }
}

// Stages one particle's 27 node values {wt (u + a Ac0 + b Ac1 + c Ac2), wt} (.w: the bare weight) into its tile row.
// (Carrying the value as the fp32x2 pairs (x, y), (z, 1) -- FMUL2 / FFMA2 -- was measured slower: the register-pair moves
// eat the saved issue slots, 128.1 vs 126.0 us per launch, profiles/r02_run22.sh.)
__device__ __forceinline__ void p2g_stage_row(const Stencil& st, const float u[3], const float Ac[3][3], float lw,
                                              float4* __restrict__ myrow) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float wa = st.w[a][0] * lw;
    const float ua[3] = {u[0] + (float)a * Ac[0][0], u[1] + (float)a * Ac[0][1], u[2] + (float)a * Ac[0][2]};
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const float wab = wa * st.w[b][1];
      const float uab[3] = {ua[0] + (float)b * Ac[1][0], ua[1] + (float)b * Ac[1][1], ua[2] + (float)b * Ac[1][2]};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float wt = wab * st.w[c][2];
        myrow[a * 9 + b * 3 + c] = make_float4(wt * (uab[0] + (float)c * Ac[2][0]), wt * (uab[1] + (float)c * Ac[2][1]),
                                               wt * (uab[2] + (float)c * Ac[2][2]), wt);   // .w: bare weight
      }
    }
  }
}

constexpr int P2GW_BLOCK = 128;   // 4 independent warps per CTA (the CTA is only the unit shared memory is carved in)
constexpr int P2GW_MIN_BLOCKS = 4;  // resident CTAs per SM = what the 14 KB staging tile per warp leaves room for
// One tile (32 particles) per warp, one CTA per 4 tiles: the A/B partner of k_p2g_pers (ud_tuning_set("pers", 0)).
template <bool DET, bool LIQ>
__global__ void __launch_bounds__(P2GW_BLOCK, P2GW_MIN_BLOCKS)
k_p2g_warp(MpmConst k, const float* ps_in, float* ps_out, float4* __restrict__ grid,
           const float* __restrict__ mu_s, const float* __restrict__ la_s, const int32_t* __restrict__ mat_s,
           const float* __restrict__ h_s, const float* __restrict__ vt_in, float* __restrict__ vt_out,
           float* __restrict__ svd_out, int vt_svd_i, BlkList bl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* wmem = smem_raw + (threadIdx.x >> 5) * WARP_TILE_BYTES;
  float4* tile = reinterpret_cast<float4*>(wmem);                                                  // [32][27]
  int2* seg = reinterpret_cast<int2*>(wmem + sizeof(float4) * 32 * WARP_TILE_NODES);               // segment table
  UD_PARTICLE_INDEX(k, env, g);
  if (slot_ - (int)(threadIdx.x & 31) >= k.n) return;   // warp-uniform: tiles beyond the env's last particle
  {
    Stencil st;
    float u[3], Ac[3][3];
    p2g_front<LIQ>(k, env, g, gp, true, live_, ps_in, ps_out, mu_s, la_s, mat_s, h_s, vt_in, vt_out, svd_out, vt_svd_i != 0, st, u, Ac);
    const WarpGroup wg = warp_group(k, live_ ? (base_key(st.base) & ~DEAD_KEY) : DEAD_KEY, st.base, live_, seg);
    float4* genv = grid + (size_t)env * k.G * (DET ? 2 : 1);   // DET: the int64 accumulator array (32 B per cell)
    float4* myrow = tile + wg.row * WARP_TILE_NODES;
    const float lw = live_ ? 1.f : 0.f;
    __syncwarp();   // (persistent kernel: every lane has left the previous tile's flush)
    p2g_stage_row(st, u, Ac, lw, myrow);
    __syncwarp();
    warp_flush<false, DET>(k, tile, seg, wg, st.base, genv, k.p_mass);
    warp_mark_blocks(k, wg, st.base, live_, bl, env);
  }
}

// Persistent variant: one CTA slot per resident CTA, every warp walks the tiles w, w + W, w + 2W, ... and issues the
// loads of its NEXT tile before it starts computing the current one, so a tile's DRAM latency (12 % of the one-tile
// kernel's stall samples sat on the first use of x) is hidden behind a whole tile of arithmetic.
template <bool DET, bool LIQ>
__global__ void __launch_bounds__(P2GW_BLOCK, P2GW_MIN_BLOCKS)
k_p2g_pers(MpmConst k, const float* ps_in, float* ps_out, float4* __restrict__ grid,
           const float* __restrict__ mu_s, const float* __restrict__ la_s, const int32_t* __restrict__ mat_s,
           const float* __restrict__ h_s, const float* __restrict__ vt_in, float* __restrict__ vt_out,
           float* __restrict__ svd_out, int vt_svd_i, BlkList bl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* wmem = smem_raw + (threadIdx.x >> 5) * WARP_TILE_BYTES;
  float4* tile = reinterpret_cast<float4*>(wmem);                                                  // [32][27]
  int2* seg = reinterpret_cast<int2*>(wmem + sizeof(float4) * 32 * WARP_TILE_NODES);               // segment table
  const TileWalk w = tile_walk(k);
  if (w.t >= w.ntiles) return;
  const bool warm = vt_in != nullptr, vt_svd = vt_svd_i != 0;
  P2gIn nx;
  TilePos pos = tile_first(w);
  {
    int env, g, gp;
    bool live;
    tile_locate(k, w, pos, env, g, gp, live);
    p2g_issue_loads(env, g, gp, ps_in, mu_s, la_s, mat_s, h_s, vt_in, vt_svd, nx);
  }
  for (; pos.t < w.ntiles;) {
    int env, g, gp;
    bool live_;
    tile_locate(k, w, pos, env, g, gp, live_);
    const P2gIn cur = nx;
    pos = tile_next(w, pos);
    if (pos.t < w.ntiles) {
      int e2, g2, gp2;
      bool l2;
      tile_locate(k, w, pos, e2, g2, gp2, l2);
      p2g_issue_loads(e2, g2, gp2, ps_in, mu_s, la_s, mat_s, h_s, vt_in, vt_svd, nx);
    }
    Stencil st;
    float u[3], Ac[3][3];
    p2g_front_loaded<LIQ>(k, gp, true, live_, cur, warm, ps_out, vt_out, svd_out, st, u, Ac);
    const WarpGroup wg = warp_group(k, live_ ? (base_key(st.base) & ~DEAD_KEY) : DEAD_KEY, st.base, live_, seg);
    float4* genv = grid + (size_t)env * k.G * (DET ? 2 : 1);   // DET: the int64 accumulator array (32 B per cell)
    float4* myrow = tile + wg.row * WARP_TILE_NODES;
    const float lw = live_ ? 1.f : 0.f;
    __syncwarp();   // (persistent kernel: every lane has left the previous tile's flush)
    p2g_stage_row(st, u, Ac, lw, myrow);
    __syncwarp();
    warp_flush<false, DET>(k, tile, seg, wg, st.base, genv, k.p_mass);
    warp_mark_blocks(k, wg, st.base, live_, bl, env);
  }
}

static int g_pers = 1;   // 1: persistent P2G with next-tile prefetch; 0: one tile per warp
int tuning_pers(int v) {
  int o = g_pers;
  if (v == 0 || v == 1) g_pers = v;
  return o;
}

static int g_warp_nw = 1;   // 1: warp-local staged-scatter kernels; 0 = the round-1 CTA-staged kernels (A/B partner)
int tuning_warp(int v) {
  int o = g_warp_nw;
  if (v == 0 || v == 1) g_warp_nw = v;
  return o;
}

static BlkList blk_list_of(const MpmConst& k, const MpmWs& ws, int substep) {
  const size_t total = (size_t)k.B * k.nbx * k.nby * k.nbz;
  BlkList bl = {ws.blk_flag, ws.blk_list + (size_t)(substep % ws.blk_nbuf) * total, ws.blk_count + 2 * substep, substep + 1};
  return bl;
}

template <bool DET, bool LIQ>
static void launch_p2g_warp(const MpmConst& k, const float* ps_in, float* ps_out, float4* grid, const float* mu_s,
                            const float* la_s, const float* vt_in, float* vt_out, float* svd_out, int substep,
                            const MpmWs& ws, cudaStream_t st, bool vt_svd) {
  const size_t smem = WARP_TILE_BYTES * (P2GW_BLOCK / 32);
  // per-DEVICE attribute (one host thread per device under pmap): set on every launch
  if (g_pers) {
    cudaFuncSetAttribute(k_p2g_pers<DET, LIQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_p2g_pers<DET, LIQ><<<persistent_ctas(k, P2GW_BLOCK / 32, P2GW_MIN_BLOCKS), P2GW_BLOCK, smem, st>>>(
        k, ps_in, ps_out, grid, mu_s, la_s, ws.mat_s, ws.h_s, vt_in, vt_out, svd_out, (int)vt_svd, blk_list_of(k, ws, substep));
    return;
  }
  cudaFuncSetAttribute(k_p2g_warp<DET, LIQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_p2g_warp<DET, LIQ><<<pgrid(k, P2GW_BLOCK), P2GW_BLOCK, smem, st>>>(k, ps_in, ps_out, grid, mu_s, la_s, ws.mat_s, ws.h_s, vt_in,
                                                                  vt_out, svd_out, (int)vt_svd, blk_list_of(k, ws, substep));
}

// true: the P2G kernels in use append the touched blocks to the substep's list themselves (no k_blk_compact pass)
bool p2g_lists_blocks() { return tuning_stage() && g_warp_nw; }

void launch_p2g(const MpmConst& k, const float* ps_in, float* ps_out, float4* grid, const float* mu_s,
                const float* la_s, const float* vt_in, float* vt_out, float* svd_out, int substep, const MpmWs& ws,
                cudaStream_t st, bool vt_in_is_vt) {
  KScope ks_(KC_P2G, st);
  const bool vt_svd = svd_out != nullptr && !vt_in_is_vt;   // layout of vt_in: previous substep's SVD tile / V^T buffer
  if (tuning_stage() && g_warp_nw) {
    float4* fix = reinterpret_cast<float4*>(ws.grid_fix);
    if (fix && k.liquid_fast) launch_p2g_warp<true, true>(k, ps_in, ps_out, fix, mu_s, la_s, vt_in, vt_out, svd_out, substep, ws, st, vt_svd);
    else if (fix) launch_p2g_warp<true, false>(k, ps_in, ps_out, fix, mu_s, la_s, vt_in, vt_out, svd_out, substep, ws, st, vt_svd);
    else if (k.liquid_fast) launch_p2g_warp<false, true>(k, ps_in, ps_out, grid, mu_s, la_s, vt_in, vt_out, svd_out, substep, ws, st, vt_svd);
    else launch_p2g_warp<false, false>(k, ps_in, ps_out, grid, mu_s, la_s, vt_in, vt_out, svd_out, substep, ws, st, vt_svd);
    return;
  }
  // the attribute is per DEVICE (a process may drive several, one host thread each: SURVEY 8b): set it on every launch
  cudaFuncSetAttribute(k_p2g<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_smem_bytes<P2G_BLOCK>(4, P2G_NPH));
  cudaFuncSetAttribute(k_p2g<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_smem_bytes<P2G_BLOCK>(4, P2G_NPH));
  if (!tuning_stage() && !ws.grid_fix)
    k_p2g<2><<<pgrid(k, P2G_BLOCK), P2G_BLOCK, 0, st>>>(k, ps_in, ps_out, grid, mu_s, la_s, ws.mat_s, ws.h_s, vt_in,
                                                      vt_out, svd_out, (int)vt_svd, ws.blk_flag);
  else if (ws.grid_fix)
    k_p2g<1><<<pgrid(k, P2G_BLOCK), P2G_BLOCK, stage_smem_bytes<P2G_BLOCK>(4, P2G_NPH), st>>>(
        k, ps_in, ps_out, reinterpret_cast<float4*>(ws.grid_fix), mu_s, la_s, ws.mat_s, ws.h_s, vt_in, vt_out, svd_out,
        (int)vt_svd, ws.blk_flag);
  else
    k_p2g<0><<<pgrid(k, P2G_BLOCK), P2G_BLOCK, stage_smem_bytes<P2G_BLOCK>(4, P2G_NPH), st>>>(
        k, ps_in, ps_out, grid, mu_s, la_s, ws.mat_s, ws.h_s, vt_in, vt_out, svd_out, (int)vt_svd, ws.blk_flag);
}

// ------------------------------------------------------------------------------------------------
// G2P (g2p_micro, :196-221) + advection (:326).  Out-of-range nodes clamp (JAX gather rule).
// Rows of C' of original particles 0..2 are kept for the J update quirk (:327).
// ------------------------------------------------------------------------------------------------
// Per-warp node tile.  The 32 particles of a warp sit in a few cells (sorted order), so the warp fetches the 27 grid
// values of each DISTINCT base cell once (<= G2P_TILE_CELLS cells; all loads are issued before the first one is
// consumed: one L2 round trip, no block barrier) into `tile` [G2P_TILE_CELLS][32] and every lane reads its stencil from
// shared memory; warps spanning more cells gather from L1/L2.  (Sending an interior cell as one shuffle of its linear
// index instead of three + the index rules was measured slower in k_g2p, 46.0 -> 49.5 us: spills at its 64 registers.)  CLAMP: gather rule (clamped indices); otherwise the
// scatter rule's transpose (dropped nodes read as zero).  Returns false (warp-uniform) when the warp spans more cells
// than the tile holds; *gid = index of my cell in the tile.  All lanes of the warp must call it.
template <bool CLAMP>
__device__ __forceinline__ bool warp_tile_fill(const MpmConst& k, const float4* __restrict__ genv, const int base[3], bool live,
                                               float4* __restrict__ tile, int* gid_out) {
  // tile layout: [cell][32] float4, node j = a*9 + b*3 + c of cell ge at tile[ge * 32 + j]: lane j fetches node j of
  // every cell, so the node offsets (a, b, c) are per-lane constants (no per-element index arithmetic)
  const int lane = threadIdx.x & 31;
  const int key = live ? base_key(base) : 0x7fffffff - lane;   // dead lanes: singleton groups no base can collide with
  const unsigned mm = __match_any_sync(0xffffffffu, key);
  const int leadlane = __ffs(mm) - 1;
  const unsigned lb = __ballot_sync(0xffffffffu, lane == leadlane && live);
  const int ngroups = __popc(lb);
  *gid_out = __popc(lb & ((1u << leadlane) - 1u));
  const bool tiled = ngroups <= G2P_TILE_CELLS && ngroups > 0;    // warp-uniform
  if (tiled) {
    const int jn = lane < 27 ? lane : 0;
    const int a = jn / 9, b = (jn / 3) % 3, c = jn % 3;
    float4 tmp[G2P_TILE_CELLS];
    unsigned bits = lb;
#pragma unroll
    for (int ge = 0; ge < G2P_TILE_CELLS; ++ge) {
      tmp[ge] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ge < ngroups) {   // warp-uniform
        const int L = __ffs(bits) - 1;
        bits &= bits - 1;
        const int bx = __shfl_sync(0xffffffffu, base[0], L), by = __shfl_sync(0xffffffffu, base[1], L),
                  bz = __shfl_sync(0xffffffffu, base[2], L);
        int ix, iy, iz;
        if (CLAMP) {
          ix = idx_gather(bx + a, k.rx);
          iy = idx_gather(by + b, k.ry);
          iz = idx_gather(bz + c, k.rz);
        } else {
          ix = idx_scatter(bx + a, k.rx);
          iy = idx_scatter(by + b, k.ry);
          iz = idx_scatter(bz + c, k.rz);
        }
        if (lane < 27 && (CLAMP || (ix | iy | iz) >= 0)) tmp[ge] = __ldg(&genv[(ix * k.ry + iy) * k.rz + iz]);
      }
    }
#pragma unroll
    for (int ge = 0; ge < G2P_TILE_CELLS; ++ge)
      if (ge < ngroups && lane < 27) tile[ge * NT_STRIDE + lane] = tmp[ge];
    __syncwarp();
  }
  return tiled;
}

// nv = sum wt g,  nC = 4 inv_dx sum wt g (x) (off - fx), contracted axis by axis (z, then y, then x) instead of node
// by node: with o' = off - 1 in {-1, 0, 1} and f' = fx - 1 in [-0.5, 0.5),
//   nC_ij = 4 inv_dx (X_j[i] - f'_j nv_i),   X_j = sum wt g o'_j
// and every partial sum over c (then b) is shared by the four quantities: ~260 instead of ~430 FP instructions.
__device__ __forceinline__ void g2p_gather(const MpmConst& k, const float4* __restrict__ genv, const Stencil& st, bool live,
                                           float4* __restrict__ tile, float nv[3], Mat3& nC) {
  int gid;
  const bool tiled = warp_tile_fill<true>(k, genv, st.base, live, tile, &gid);
#pragma unroll
  for (int c = 0; c < 3; ++c) nv[c] = 0.f;
  nC = mat_zero();
  if (!live) return;
  const float4* my_tile = tile + gid * NT_STRIDE;
  float Xa[3] = {0.f, 0.f, 0.f}, Xb[3] = {0.f, 0.f, 0.f}, Xc[3] = {0.f, 0.f, 0.f};
  auto nodes = [&](auto tiled_tag) {
    constexpr bool TILED = decltype(tiled_tag)::value;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ix = TILED ? 0 : idx_gather(st.base[0] + a, k.rx);
      float U0[3] = {0.f, 0.f, 0.f}, Ub[3] = {0.f, 0.f, 0.f}, Uc[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int iy = TILED ? 0 : idx_gather(st.base[1] + b, k.ry);
        float4 gv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (TILED) {
            gv[c] = my_tile[a * 9 + b * 3 + c];
          } else {
            const int iz = idx_gather(st.base[2] + c, k.rz);
            gv[c] = __ldg(&genv[(ix * k.ry + iy) * k.rz + iz]);
          }
        }
        const float g0[3] = {gv[0].x, gv[0].y, gv[0].z}, g1[3] = {gv[1].x, gv[1].y, gv[1].z}, g2[3] = {gv[2].x, gv[2].y, gv[2].z};
        const float wy = st.w[b][1];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float p0 = st.w[0][2] * g0[i], p2 = st.w[2][2] * g2[i];
          const float T0 = (p0 + st.w[1][2] * g1[i]) + p2;   // sum_c w_c g
          const float T2 = p2 - p0;                            // sum_c w_c o'_c g
          const float q = wy * T0;
          U0[i] += q;
          if (b == 0) Ub[i] -= q;
          if (b == 2) Ub[i] += q;
          Uc[i] += wy * T2;
        }
      }
      const float wx = st.w[a][0];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float q = wx * U0[i];
        nv[i] += q;
        if (a == 0) Xa[i] -= q;
        if (a == 2) Xa[i] += q;
        Xb[i] += wx * Ub[i];
        Xc[i] += wx * Uc[i];
      }
    }
  };
  if (tiled) nodes(std::true_type{});
  else nodes(std::false_type{});
  const float c4 = 4.f * k.inv_dx;   // the factor is applied once, after the 27-node sum
  const float f0 = st.fx[0] - 1.f, f1 = st.fx[1] - 1.f, f2 = st.fx[2] - 1.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    nC(i, 0) = c4 * (Xa[i] - f0 * nv[i]);
    nC(i, 1) = c4 * (Xb[i] - f1 * nv[i]);
    nC(i, 2) = c4 * (Xc[i] - f2 * nv[i]);
  }
}

#ifndef UD_G2P_MINB
#define UD_G2P_MINB 8
#endif
__global__ void __launch_bounds__(UD_BLOCK, UD_G2P_MINB)
k_g2p(MpmConst k, const float* ps_in, float* ps_out, const float4* __restrict__ grid,
      const int32_t* __restrict__ perm, float* __restrict__ jrows, int substep) {
  __shared__ float4 wtile[UD_BLOCK / 32][G2P_TILE_CELLS * 32];
  UD_PARTICLE_INDEX_REV(k, env, g);
  if (slot_ - (int)(threadIdx.x & 31) >= k.n) return;   // warp-uniform
  const int p = perm[g];   // needed only at the end: issued with the first load
  float x[3];
  load_comps<PS_X, 3, PS_NQ>(ps_in, gp, x);
  Stencil st;
  make_stencil(x, k.inv_dx, st);
  float nv[3];
  Mat3 nC;
  g2p_gather(k, grid + (size_t)env * k.G, st, live_, wtile[threadIdx.x >> 5], nv, nC);
  // padding lanes of the env's last tile (nv = nC = 0) stay a harmless particle at the origin
  float outv[15];
#pragma unroll
  for (int c = 0; c < 9; ++c) outv[PS_C + c] = nC.m[c];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    outv[PS_X + c] = live_ ? x[c] + k.dt * nv[c] : 0.f;
    outv[PS_V + c] = nv[c];
  }
  store_comps<0, 15, PS_NQ>(ps_out, gp, outv);
  if (live_ && p < 3) {
    float* jr = jrows + ((size_t)env * k.S + substep) * 9 + p * 3;
    jr[0] = nC(p, 0);
    jr[1] = nC(p, 1);
    jr[2] = nC(p, 2);
  }
}

void launch_g2p(const MpmConst& k, const float* ps_in, float* ps_out, const float4* grid, int substep,
                const MpmWs& ws, cudaStream_t st) {
  KScope ks_(KC_G2P, st);
  k_g2p<<<pgrid(k, UD_BLOCK), UD_BLOCK, 0, st>>>(k, ps_in, ps_out, grid, ws.perm, ws.jrows, substep);
}

// Stores one AoS leaf with NC floats per particle for the warp's consecutive particles [o0, o0 + nlive): the lanes'
// values go through a per-warp shared-memory tile so that every store instruction writes 32 CONSECUTIVE floats
// (one 128-byte line) instead of 32 floats NC apart (NC lines): 3 / 9 times fewer sectors per request.
template <int NC>
__device__ __forceinline__ void warp_store_aos(float* __restrict__ dst, size_t o0, int nlive, const float* vals,
                                               float* __restrict__ sm) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NC; ++c) sm[lane * NC + c] = vals[c];   // NC is odd (3, 9): conflict-free
  __syncwarp();
  float* out = dst + o0 * NC;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int e = j * 32 + lane;
    if (e < nlive * NC) out[e] = sm[e];
  }
}

// sorted tiles -> AoS outputs; J' = J * prod_f (1 + dt * trace-quirk_f), sequentially as the reference.
// One thread per ORIGINAL particle, which gathers its sorted slot through the inverse permutation; the AoS stores go
// through warp_store_aos (round 1 scattered the stores: 160 us; gather + strided stores: 139 us).
__global__ void __launch_bounds__(256)
k_unsort_state(MpmConst k, const float* __restrict__ ps, const float* __restrict__ J_in,
               const int32_t* __restrict__ inv_perm, const float* __restrict__ jrows,
               float* __restrict__ x, float* __restrict__ v, float* __restrict__ C,
               float* __restrict__ F, float* __restrict__ J) {
  __shared__ float sm[256 / 32][32 * 9];
  __shared__ float jfac[128];
  UD_PARTICLE_INDEX(k, env, g);
  // the J factor of every substep is the same for all particles of the env (the trace quirk reads rows of particles
  // 0..2): formed once per CTA, applied sequentially per particle as before
  const int nr = min(3, k.n);
  if (threadIdx.x < 128 && (int)threadIdx.x < k.S && k.S <= 128) {   // S <= 128 in every shipped scene
    const float* jr = jrows + ((size_t)env * k.S + threadIdx.x) * 9;
    float t[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < nr; ++i) {
      t[0] += jr[3 * i];
      t[1] += jr[3 * i + 1];
      t[2] += jr[3 * i + 2];
    }
    jfac[threadIdx.x] = 1.f + k.dt * ((t[0] + t[1]) + t[2]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int slot0 = slot_ - lane;
  if (slot0 >= k.n) return;                 // warp-uniform
  const int nlive = min(32, k.n - slot0);
  const size_t o0 = (size_t)env * k.n + slot0;
  const int sp = env * k.n_pad + (live_ ? inv_perm[g] : 0);   // sorted slot of the original particle
  float st[PS_NCOMP];
  load_comps<0, PS_NCOMP, PS_NQ>(ps, sp, st);
  float* wsm = sm[threadIdx.x >> 5];
  warp_store_aos<3>(x, o0, nlive, st + PS_X, wsm);
  warp_store_aos<3>(v, o0, nlive, st + PS_V, wsm);
  warp_store_aos<9>(C, o0, nlive, st + PS_C, wsm);
  warp_store_aos<9>(F, o0, nlive, st + PS_F, wsm);
  if (!live_) return;
  float j = nan_to_num(J_in[g]);
  if (k.S <= 128) {
    for (int f = 0; f < k.S; ++f) j = j * jfac[f];
  } else {   // longer steps than any shipped scene: straight from global memory
    for (int f = 0; f < k.S; ++f) {
      const float* jr = jrows + ((size_t)env * k.S + f) * 9;
      float t[3] = {0.f, 0.f, 0.f};
      for (int i = 0; i < nr; ++i) {
        t[0] += jr[3 * i];
        t[1] += jr[3 * i + 1];
        t[2] += jr[3 * i + 2];
      }
      j = j * (1.f + k.dt * ((t[0] + t[1]) + t[2]));
    }
  }
  J[g] = j;
}

void launch_unsort_state(const MpmConst& k, const float* ps_slot, const float* J_in, const MpmWs& ws,
                         ud_mpm_state* out, cudaStream_t st) {
  KScope ks_(KC_UNSORT, st);
  k_unsort_state<<<pgrid(k, 256), 256, 0, st>>>(k, ps_slot, J_in, ws.inv_perm, ws.jrows, out->x, out->v, out->C,
                                                  out->F, out->J);
}

// ================================================================================================
// Adjoint.  Cotangents live in ws.gs in the same sorted SoA layout as the state.  Per substep, in
// reverse order:  k_g2p_bwd (scatter of the grid-velocity cotangent = G2P^T + advection^T),
// k_grid_bwd (mpm_grid.cu), k_p2g_bwd (gather = P2G^T, then constitutive / SVD / F-update reverse).
// ================================================================================================
__global__ void k_gather_cot(MpmConst k, const float* __restrict__ gx, const float* __restrict__ gv,
                             const float* __restrict__ gC, const float* __restrict__ gF,
                             const int32_t* __restrict__ perm, float* __restrict__ gs) {
  UD_PARTICLE_INDEX(k, env, g);
  if (slot_ >= k.n_pad) return;
  float st[PS_NCOMP];
#pragma unroll
  for (int c = 0; c < PS_NCOMP; ++c) st[c] = 0.f;   // padding slots carry zero cotangents
  if (live_) {
    int p = perm[g];
    size_t o = (size_t)env * k.n + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) st[PS_X + c] = gx ? gx[3 * o + c] : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) st[PS_V + c] = gv ? gv[3 * o + c] : 0.f;
#pragma unroll
    for (int c = 0; c < 9; ++c) st[PS_C + c] = gC ? gC[9 * o + c] : 0.f;
#pragma unroll
    for (int c = 0; c < 9; ++c) st[PS_F + c] = gF ? gF[9 * o + c] : 0.f;
  }
  store_comps<0, PS_NCOMP, PS_NQ>(gs, gp, st);
}

void launch_gather_cot(const MpmConst& k, const ud_mpm_state* gout, const MpmWs& ws, cudaStream_t st) {
  KScope ks_(KC_GATHER, st);
  k_gather_cot<<<pgrid(k, 256), 256, 0, st>>>(k, gout->x, gout->v, gout->C, gout->F, ws.perm, ws.gs);
}

// loads of G2P^T: the particle's position and the cotangents of (x', v', C') it produced
__device__ __forceinline__ void g2pb_load(const MpmConst& k, const float* __restrict__ ps_in, const float* __restrict__ gs,
                                          int gp, float x[3], float gxo[3], float gvt[3], Mat3& gC) {
  load_comps<PS_X, 3, PS_NQ>(ps_in, gp, x);
  float g15[15];
  load_comps<0, 15, PS_NQ>(gs, gp, g15);
#pragma unroll
  for (int c = 0; c < 3; ++c) gxo[c] = g15[PS_X + c];
#pragma unroll
  for (int c = 0; c < 3; ++c) gvt[c] = g15[PS_V + c] + k.dt * gxo[c];
#pragma unroll
  for (int c = 0; c < 9; ++c) gC.m[c] = g15[PS_C + c];
}

// G2P^T: v' = sum wt g ; C' = sum 4 inv_dx wt g (x) d ; x' = x + dt v'   (d = off - fx)
// Scatters the cotangent of the updated grid velocity (clamped index = transpose of the clamping
// gather) and leaves the partial position cotangent gx' + inv_dx * gfx in gs.X.
__global__ void __launch_bounds__(G2PB_BLOCK)
k_g2p_bwd(MpmConst k, const float* __restrict__ ps_in, const float4* __restrict__ grid_out,
          float* __restrict__ gs, float4* __restrict__ ggrid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sv = reinterpret_cast<float*>(smem_raw);                               // [81][STG_PAD]
  constexpr int STG_PAD = stg_pad<G2PB_BLOCK>();
  StageMeta<G2PB_BLOCK>& meta = *reinterpret_cast<StageMeta<G2PB_BLOCK>*>(sv + G2PB_NPH * 3 * STG_PAD);
  float4* tile = reinterpret_cast<float4*>(smem_raw + g2pb_tile_offset());  // [G2PB_TILE_RUNS][27]
  UD_PARTICLE_INDEX(k, env, g);
  float x[3], gxo[3], gvt[3];
  Mat3 gC;
  g2pb_load(k, ps_in, gs, gp, x, gxo, gvt, gC);
  Stencil st;
  make_stencil(x, k.inv_dx, st);
  int my_run = 0;
  const int row = stage_group<true, G2PB_BLOCK>(k, meta, live_ ? (base_key(st.base) & ~DEAD_KEY) : DEAD_KEY, st.base,
                                                nullptr, &my_run);
  const float lw = live_ ? 1.f : 0.f;
  const float4* genv = grid_out + (size_t)env * k.G;
  // The CTA's particles sit in a handful of cells: the 27 grid velocities of every segment are fetched ONCE into
  // shared memory (coalesced over the cell table) and every particle of the segment reads them from there.
  __syncthreads();   // publishes the cell table
  const bool tiled = meta.n_runs <= G2PB_TILE_RUNS;   // block-uniform
  if (tiled) {
    for (int e = threadIdx.x; e < meta.n_runs * 27; e += G2PB_BLOCK) {
      const int cell = meta.cell[e / 27][e % 27];
      tile[e] = cell >= 0 ? __ldg(&genv[cell]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
  }
  const float4* my_tile = tile + my_run * 27;
  const float c4 = 4.f * k.inv_dx;
  // r(a,b,c) = gv' + 4 inv_dx gC' (off - fx) = r0 + a K0 + b K1 + c K2   (K_j = 4 inv_dx * column j of gC')
  float K[3][3], r0[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) K[j][i] = c4 * gC(i, j);
    r0[i] = gvt[i] - (K[0][i] * st.fx[0] + K[1][i] * st.fx[1] + K[2][i] * st.fx[2]);
  }
  // accumulators: Wg = sum wt g ; the weight cotangent gwt = g . r is contracted hierarchically with (w, dw)
  float Wg[3] = {0.f, 0.f, 0.f}, gfx[3] = {0.f, 0.f, 0.f};
  auto nodes = [&](auto tiled_tag) {
    constexpr bool TILED = decltype(tiled_tag)::value;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ix = TILED ? 0 : idx_gather(st.base[0] + a, k.rx);
      const float wa = st.w[a][0] * lw;
      float ra[3] = {r0[0] + (float)a * K[0][0], r0[1] + (float)a * K[0][1], r0[2] + (float)a * K[0][2]};
      float P1 = 0.f, P2 = 0.f, Q1 = 0.f;  // sum_b (sum_c gwt w_c) w_b, ... dw_b, sum_b (sum_c gwt dw_c) w_b
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int iy = TILED ? 0 : idx_gather(st.base[1] + b, k.ry);
        const float wab = wa * st.w[b][1];
        float rab[3] = {ra[0] + (float)b * K[1][0], ra[1] + (float)b * K[1][1], ra[2] + (float)b * K[1][2]};
        float P = 0.f, Q = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float wt = wab * st.w[c][2];
          float4 gv;
          if (TILED) {
            gv = my_tile[a * 9 + b * 3 + c];
          } else {
            const int iz = idx_gather(st.base[2] + c, k.rz);
            gv = __ldg(&genv[(ix * k.ry + iy) * k.rz + iz]);
          }
          const float r[3] = {rab[0] + (float)c * K[2][0], rab[1] + (float)c * K[2][1], rab[2] + (float)c * K[2][2]};
          float* dst = sv + ((a * 9 + b * 3 + c) * 3) * STG_PAD + row;
          dst[0] = wt * r[0];
          dst[STG_PAD] = wt * r[1];
          dst[2 * STG_PAD] = wt * r[2];
          const float gwt = gv.x * r[0] + gv.y * r[1] + gv.z * r[2];
          Wg[0] += wt * gv.x;
          Wg[1] += wt * gv.y;
          Wg[2] += wt * gv.z;
          P += gwt * st.w[c][2];
          Q += gwt * st.dw[c][2];
        }
        P1 += P * st.w[b][1];
        P2 += P * st.dw[b][1];
        Q1 += Q * st.w[b][1];
      }
      gfx[0] += P1 * st.dw[a][0];
      gfx[1] += P2 * st.w[a][0];
      gfx[2] += Q1 * st.w[a][0];
    }
  };
  if (tiled) nodes(std::true_type{});
  else nodes(std::false_type{});
  if (live_) {
    // fx enters d = off - fx with a minus sign: gfx_j -= sum_n wt (K_j . g) = K_j . Wg
#pragma unroll
    float ox[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      gfx[d] -= K[d][0] * Wg[0] + K[d][1] * Wg[1] + K[d][2] * Wg[2];
      ox[d] = gxo[d] + k.inv_dx * gfx[d];
    }
    store_comps<PS_X, 3, PS_NQ>(gs, gp, ox);
  }
  __syncthreads();
  // transpose of the clamping gather: clamped target index
  stage_flush<3, true, false, G2PB_BLOCK>(k, sv, meta, ggrid + (size_t)env * k.G, 0, 27);
}

// Warp-local G2P^T (round 2): the same arithmetic as k_g2p_bwd, but a warp owns its 32 particles end to end --
// warp_group() for the segments, the 27 grid velocities of each distinct cell fetched once per WARP into a small
// tile (all loads issued before the first use), the cotangent scatter staged as float4 rows and flushed by
// warp_flush() with clamped target indices (transpose of the clamping gather).  No block barrier.
constexpr int G2PBW_BLOCK = 64;
constexpr int G2PBW_CELLS = 4;   // distinct cells per warp whose grid velocities are tiled; more -> gather from L1/L2
constexpr size_t G2PBW_WARP_BYTES = WARP_TILE_BYTES + sizeof(float4) * G2PBW_CELLS * 32;
constexpr int G2PBW_MIN_BLOCKS = 7;   // resident 2-warp CTAs per SM that the staging tiles leave room for
__global__ void __launch_bounds__(G2PBW_BLOCK, G2PBW_MIN_BLOCKS)
k_g2p_bwd_warp(MpmConst k, const float* __restrict__ ps_in, const float4* __restrict__ grid_out,
               float* __restrict__ gs, float4* __restrict__ ggrid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* wbase = smem_raw + (threadIdx.x >> 5) * G2PBW_WARP_BYTES;
  float4* tile = reinterpret_cast<float4*>(wbase);                                            // [32][27] staged cotangents
  int2* seg = reinterpret_cast<int2*>(wbase + sizeof(float4) * 32 * WARP_TILE_NODES);        // segment table
  float4* vtile = reinterpret_cast<float4*>(wbase + WARP_TILE_BYTES);                         // [G2PBW_CELLS][32] grid velocities
  UD_PARTICLE_INDEX(k, env, g);
  const int lane = threadIdx.x & 31;
  if (slot_ - lane >= k.n) return;   // warp-uniform
  float x[3], gxo[3], gvt[3];
  Mat3 gC;
  g2pb_load(k, ps_in, gs, gp, x, gxo, gvt, gC);
  ahead_prefetch<0, 0, PS_NQ>(k, ps_in, gp);
  ahead_prefetch<0, 3, PS_NQ>(k, gs, gp);
  Stencil st;
  make_stencil(x, k.inv_dx, st);
  const WarpGroup wg = warp_group(k, live_ ? (base_key(st.base) & ~DEAD_KEY) : DEAD_KEY, st.base, live_, seg);
  const float4* genv = grid_out + (size_t)env * k.G;
  const int nseg = __popc(wg.lb);
  const int leadlane = __ffs(wg.mm) - 1;
  const int my_seg = __popc(wg.lb & ((1u << leadlane) - 1u));
  const bool tiled = nseg <= G2PBW_CELLS;   // warp-uniform
  if (tiled) {   // vtile layout [cell][32]: lane j fetches node j of every cell (per-lane constant node offsets)
    const int jn = lane < 27 ? lane : 0;
    const int ja = jn / 9, jb = (jn / 3) % 3, jc = jn % 3;
    float4 tmp[G2PBW_CELLS];
    unsigned bits = wg.lb;
#pragma unroll
    for (int ge = 0; ge < G2PBW_CELLS; ++ge) {
      if (ge < nseg) {   // warp-uniform
        const int L = __ffs(bits) - 1;
        bits &= bits - 1;
        const int b0 = __shfl_sync(FULL, st.base[0], L), b1 = __shfl_sync(FULL, st.base[1], L),
                  b2 = __shfl_sync(FULL, st.base[2], L);
        const int ix = idx_gather(b0 + ja, k.rx), iy = idx_gather(b1 + jb, k.ry), iz = idx_gather(b2 + jc, k.rz);
        tmp[ge] = __ldg(&genv[(ix * k.ry + iy) * k.rz + iz]);
      }
    }
#pragma unroll
    for (int ge = 0; ge < G2PBW_CELLS; ++ge)
      if (ge < nseg && lane < 27) vtile[ge * NT_STRIDE + lane] = tmp[ge];
    __syncwarp();
  }
  const float4* my_tile = vtile + my_seg * NT_STRIDE;
  const float lw = live_ ? 1.f : 0.f;
  const float c4 = 4.f * k.inv_dx;
  // r(a,b,c) = gv' + 4 inv_dx gC' (off - fx) = r0 + a K0 + b K1 + c K2   (K_j = 4 inv_dx * column j of gC')
  float K[3][3], r0[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) K[j][i] = c4 * gC(i, j);
    r0[i] = gvt[i] - (K[0][i] * st.fx[0] + K[1][i] * st.fx[1] + K[2][i] * st.fx[2]);
  }
  float Wg[3] = {0.f, 0.f, 0.f}, gfx[3] = {0.f, 0.f, 0.f};
  float4* myrow = tile + wg.row * WARP_TILE_NODES;
  float4* ggenv = ggrid + (size_t)env * k.G;
  auto nodes = [&](auto tiled_tag) {
    constexpr bool TILED = decltype(tiled_tag)::value;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ix = TILED ? 0 : idx_gather(st.base[0] + a, k.rx);
      const float wa = st.w[a][0] * lw;
      float ra[3] = {r0[0] + (float)a * K[0][0], r0[1] + (float)a * K[0][1], r0[2] + (float)a * K[0][2]};
      float P1 = 0.f, P2 = 0.f, Q1 = 0.f;  // sum_b (sum_c gwt w_c) w_b, ... dw_b, sum_b (sum_c gwt dw_c) w_b
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int iy = TILED ? 0 : idx_gather(st.base[1] + b, k.ry);
        const float wab = wa * st.w[b][1];
        float rab[3] = {ra[0] + (float)b * K[1][0], ra[1] + (float)b * K[1][1], ra[2] + (float)b * K[1][2]};
        float P = 0.f, Q = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float wt = wab * st.w[c][2];
          float4 gv;
          if (TILED) {
            gv = my_tile[a * 9 + b * 3 + c];
          } else {
            const int iz = idx_gather(st.base[2] + c, k.rz);
            gv = __ldg(&genv[(ix * k.ry + iy) * k.rz + iz]);
          }
          const float r[3] = {rab[0] + (float)c * K[2][0], rab[1] + (float)c * K[2][1], rab[2] + (float)c * K[2][2]};
          myrow[a * 9 + b * 3 + c] = make_float4(wt * r[0], wt * r[1], wt * r[2], 0.f);
          const float gwt = gv.x * r[0] + gv.y * r[1] + gv.z * r[2];
          Wg[0] += wt * gv.x;
          Wg[1] += wt * gv.y;
          Wg[2] += wt * gv.z;
          P += gwt * st.w[c][2];
          Q += gwt * st.dw[c][2];
        }
        P1 += P * st.w[b][1];
        P2 += P * st.dw[b][1];
        Q1 += Q * st.w[b][1];
      }
      gfx[0] += P1 * st.dw[a][0];
      gfx[1] += P2 * st.w[a][0];
      gfx[2] += Q1 * st.w[a][0];
    }
  };
  if (tiled) nodes(std::true_type{});
  else nodes(std::false_type{});
  if (live_) {
    // fx enters d = off - fx with a minus sign: gfx_j -= sum_n wt (K_j . g) = K_j . Wg
    float ox[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      gfx[d] -= K[d][0] * Wg[0] + K[d][1] * Wg[1] + K[d][2] * Wg[2];
      ox[d] = gxo[d] + k.inv_dx * gfx[d];
    }
    store_comps<PS_X, 3, PS_NQ>(gs, gp, ox);
  }
  __syncwarp();
  warp_flush<true, false>(k, tile, seg, wg, st.base, ggenv);
}

int tuning_warp(int v);
static void launch_g2p_bwd_warp(const MpmConst& k, const float* ps_in, const float4* grid_out, const MpmWs& ws,
                                cudaStream_t st, float4* ggrid) {
  const size_t smem = G2PBW_WARP_BYTES * (G2PBW_BLOCK / 32);
  cudaFuncSetAttribute(k_g2p_bwd_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
  k_g2p_bwd_warp<<<pgrid(k, G2PBW_BLOCK), G2PBW_BLOCK, smem, st>>>(k, ps_in, grid_out, ws.gs, ggrid);
}

void launch_g2p_bwd(const MpmConst& k, const float* ps_in, const float4* grid_out, const MpmWs& ws,
                    cudaStream_t st, float4* ggrid) {
  KScope ks_(KC_G2P_BWD, st);
  const int nw = tuning_warp(-1);
  if (nw == 1) return launch_g2p_bwd_warp(k, ps_in, grid_out, ws, st, ggrid);
  cudaFuncSetAttribute(k_g2p_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g2pb_smem_bytes());   // per device
  k_g2p_bwd<<<pgrid(k, G2PB_BLOCK), G2PB_BLOCK, g2pb_smem_bytes(), st>>>(k, ps_in, grid_out, ws.gs, ggrid);
}

// P2G^T (gather of the cotangents of scattered momentum/mass; dropped nodes contribute nothing),
// then the reverse of stress / plasticity / SVD / F update.  Writes the cotangents of the substep's
// input x, v, C, F in place and reduces d/d(state.mu), d/d(state.lamda) per env.
// The SVD of F1 is read back from the recompute pass (svd_in) instead of being recomputed.
// Per node only q = wt*g_p, the weight cotangent gwt and a few running sums are formed; everything that
// is linear in per-particle constants (affine, fx, p_mass) is applied once after the 27-node loop:
//   gv = p_mass S            S = sum q            T_j = sum off_j q
//   gA_ij = dx (T_ij - S_i fx_j)                  gfx(direct) = -dx A^T S
//   gwt = p_mass g_m + g_p . (p_mass v + A dpos)  contracted hierarchically with (w, dw) over c, b, a
template <bool LIQ>
__global__ void __launch_bounds__(UD_BLOCK, 4)
k_p2g_bwd(MpmConst k, const float* __restrict__ ps_in, const float* __restrict__ svd_in,
          const float4* __restrict__ ggrid, float* __restrict__ gs, const float* __restrict__ mu_s,
          const float* __restrict__ la_s, const int32_t* __restrict__ mat_s, const float* __restrict__ h_s,
          float* __restrict__ g_scal, float* __restrict__ norm2) {
  // norm2 != nullptr on the step's first substep (the last one reversed): the cotangents written here are the
  // step's input cotangents, so norm_grad_state's nan_to_num + per-env sum of squares (mpm_simulator.py:389-408)
  // happen on the way out instead of in a separate pass over the 24 components
  __shared__ float4 wtile[UD_BLOCK / 32][G2P_TILE_CELLS * 32];
  const TileWalk w = tile_walk(k);
  if (w.t >= w.ntiles) return;
  // Persistent: what the tile fill needs first (x quad, material, hardness) is loaded one tile ahead into registers;
  // the other 15 quads a tile reads are prefetched into L2 one tile ahead (the kernel sits at its 128-register cap).
  float4 nxq;
  int nmat;
  float nh;
  auto issue = [&](const TilePos& tp, int prev_mat) {
    int e2, g2, gp2;
    bool l2;
    tile_locate(k, w, tp, e2, g2, gp2, l2);
    nxq = reinterpret_cast<const float4*>(ps_in)[quad_index(gp2, 0, PS_NQ)];
    nmat = mat_s[g2];
    nh = h_s[g2];
    // (one cp.async.bulk.prefetch.L2 per contiguous range from one lane instead of these 15 per-lane prefetches was
    // measured slower: 143.6 vs 141.3 us per launch)
    prefetch_quads<1, 5, PS_NQ>(ps_in, gp2);
    if (!LIQ || prev_mat != 0) prefetch_quads<0, 5, SV_NQ>(svd_in, gp2);   // (neighbouring tiles: same material, nearly always)
    prefetch_quads<0, 0, PS_NQ>(gs, gp2);
    prefetch_quads<3, 5, PS_NQ>(gs, gp2);
  };
  TilePos pos = tile_first(w);
  issue(pos, 1);
  for (; pos.t < w.ntiles;) {
  int env, g, gp;
  bool live_;
  tile_locate(k, w, pos, env, g, gp, live_);
  float gmu = 0.f, gla = 0.f, gn2 = 0.f;
  const int mat_p = nmat;
  const float h_p = nh, mu_e = mu_s[env], la_e = la_s[env];
  const float x0[3] = {nxq.x, nxq.y, nxq.z};
  pos = tile_next(w, pos);
  if (pos.t < w.ntiles) issue(pos, mat_p);
  // the cotangents of the 27 nodes of each distinct base cell of the warp, fetched once per warp (dropped nodes = 0)
  int tile_gid = 0;
  bool tiled;
  {
    Stencil s0;
    make_stencil(x0, k.inv_dx, s0);
    __syncwarp();   // every lane is done with the previous tile's node tile
    tiled = warp_tile_fill<false>(k, ggrid + (size_t)env * k.G, s0.base, live_, wtile[threadIdx.x >> 5], &tile_gid);
  }
  const float4* my_tile = wtile[threadIdx.x >> 5] + tile_gid * NT_STRIDE;
  if (live_) {
    // Phase 1: everything the 27-node gather needs is the stencil, A = dx * affine and u0.  The matrices that
    // produce them (C, F, U, s, Vt, F1, F2, D: ~80 registers) die here and are loaded again (L2 hits) for the
    // constitutive reverse after the gather, instead of staying live across it.
    Stencil st;
    float Ac[3][3], u0[3];
    const bool plastic = mat_p == 2;
    const bool liquid = LIQ && mat_p == 0;
    {
      float xvc[15], us[12];
      Mat3 C, F;
      Consti o;
      load_comps<0, 15, PS_NQ>(ps_in, gp, xvc);
#pragma unroll
      for (int c = 0; c < 9; ++c) C.m[c] = xvc[PS_C + c];
      make_stencil(xvc + PS_X, k.inv_dx, st);
      if (liquid) {   // no SVD was kept for a liquid particle (constitutive_post_liquid): J = |det F1|
        load_comps<PS_F, 9, PS_NQ>(ps_in, gp, F.m);
        liquid_affine(k, C, F, o.affine);
      } else {
        load_comps<SV_U, 12, SV_NQ>(svd_in, gp, us);
#pragma unroll
        for (int c = 0; c < 9; ++c) o.U.m[c] = us[c];
#pragma unroll
        for (int c = 0; c < 3; ++c) o.s[c] = us[9 + c];
        if (plastic) {  // the stress of a plastic particle is a function of (U, clip(s)) alone
          plastic_affine(k, C, o.U, o.s, mu_e, la_e, h_p, o.affine);
        } else {
          load_comps<PS_F, 9, PS_NQ>(ps_in, gp, F.m);
          load_comps<SV_VT, 9, SV_NQ>(svd_in, gp, o.Vt.m);
          constitutive_pre(k, C, F, mu_e, la_e, h_p, mat_p, o);
          constitutive_post(k, C, o);
        }
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Ac[j][i] = k.dx * o.affine(i, j);
        u0[i] = k.p_mass * xvc[PS_V + i] - (Ac[0][i] * st.fx[0] + Ac[1][i] * st.fx[1] + Ac[2][i] * st.fx[2]);
      }
    }
    const float4* ggenv = ggrid + (size_t)env * k.G;
    // u(a,b,c) = p_mass v + A dpos = u0 + a Ax + b Ay + c Az,  A* = dx * columns of affine (phase 1)
    float S[3] = {0.f, 0.f, 0.f}, TX[3] = {0.f, 0.f, 0.f}, TY[3] = {0.f, 0.f, 0.f}, TZ[3] = {0.f, 0.f, 0.f};
    float gfx[3] = {0.f, 0.f, 0.f};
    auto gather = [&](auto tiled_tag) {
    constexpr bool TILED = decltype(tiled_tag)::value;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ix = TILED ? 0 : idx_scatter(st.base[0] + a, k.rx);
      float ua[3] = {u0[0] + (float)a * Ac[0][0], u0[1] + (float)a * Ac[0][1], u0[2] + (float)a * Ac[0][2]};
      float Sa[3] = {0.f, 0.f, 0.f}, Ya[3] = {0.f, 0.f, 0.f}, Za[3] = {0.f, 0.f, 0.f};
      float P1 = 0.f, P2 = 0.f, Q1 = 0.f;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int iy = TILED ? 0 : idx_scatter(st.base[1] + b, k.ry);
        const float wab = st.w[a][0] * st.w[b][1];
        float uab[3] = {ua[0] + (float)b * Ac[1][0], ua[1] + (float)b * Ac[1][1], ua[2] + (float)b * Ac[1][2]};
        float Sab[3] = {0.f, 0.f, 0.f}, Zab[3] = {0.f, 0.f, 0.f};
        float P = 0.f, Q = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);  // (g_momentum, g_mass); dropped node -> 0
          if (TILED) {
            gq = my_tile[a * 9 + b * 3 + c];
          } else {
            const int iz = idx_scatter(st.base[2] + c, k.rz);
            if ((ix | iy | iz) >= 0) gq = __ldg(&ggenv[(ix * k.ry + iy) * k.rz + iz]);
          }
          const float wt = wab * st.w[c][2];
          const float q[3] = {wt * gq.x, wt * gq.y, wt * gq.z};
          const float u[3] = {uab[0] + (float)c * Ac[2][0], uab[1] + (float)c * Ac[2][1], uab[2] + (float)c * Ac[2][2]};
          const float gwt = k.p_mass * gq.w + (gq.x * u[0] + gq.y * u[1] + gq.z * u[2]);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            Sab[i] += q[i];
            Zab[i] += (float)c * q[i];
          }
          P += gwt * st.w[c][2];
          Q += gwt * st.dw[c][2];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          Sa[i] += Sab[i];
          Ya[i] += (float)b * Sab[i];
          Za[i] += Zab[i];
        }
        P1 += P * st.w[b][1];
        P2 += P * st.dw[b][1];
        Q1 += Q * st.w[b][1];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        S[i] += Sa[i];
        TX[i] += (float)a * Sa[i];
        TY[i] += Ya[i];
        TZ[i] += Za[i];
      }
      gfx[0] += P1 * st.dw[a][0];
      gfx[1] += P2 * st.w[a][0];
      gfx[2] += Q1 * st.w[a][0];
    }
    };
    if (tiled) gather(std::true_type{});
    else gather(std::false_type{});
    Mat3 gA;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gA(i, 0) = k.dx * (TX[i] - S[i] * st.fx[0]);
      gA(i, 1) = k.dx * (TY[i] - S[i] * st.fx[1]);
      gA(i, 2) = k.dx * (TZ[i] - S[i] * st.fx[2]);
    }
    // dpos = (off - fx) dx: direct fx path, gfx_j -= dx (A^T S)_j = Ac[j] . S
#pragma unroll
    for (int j = 0; j < 3; ++j) gfx[j] -= Ac[j][0] * S[0] + Ac[j][1] * S[1] + Ac[j][2] * S[2];
    // Phase 2: reload (volatile ld.global.ca: the compiler cannot keep the phase-1 registers alive, and the lines are
    // still in L1 from phase 1 -- the round-1 ld.cg reload went to L2 and held 10 % of the kernel's stall samples)
    Mat3 C, F, gF2out, gC, gF;
    Consti o;
    float gx_in[3];
    {
      float cf[18];
      load_comps<PS_C, 18, PS_NQ, true>(ps_in, gp, cf);
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        C.m[c] = cf[c];
        F.m[c] = cf[9 + c];
      }
    }
    {
      // (a liquid particle has no SVD and ignores what it loads here -- its lanes read the first tile's lines, L1 hits;
      // predicating the load instead costs the other materials' path 88 bytes of spills at the kernel's 128-register cap)
      float sv[21];
      load_comps<0, 21, SV_NQ, true>(svd_in, liquid ? (int)(threadIdx.x & 31) : gp, sv);
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        o.U.m[c] = sv[SV_U + c];
        o.Vt.m[c] = sv[SV_VT + c];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) o.s[c] = sv[SV_S + c];
    }
    load_comps<PS_X, 3, PS_NQ>(gs, gp, gx_in);
    load_comps<PS_F, 9, PS_NQ>(gs, gp, gF2out.m);
    if (liquid) {
      constitutive_bwd_liquid(k, C, F, gA, gF2out, gC, gF);
    } else if (plastic) {
      constitutive_bwd_plastic(k, C, F, o.U, o.s, o.Vt, mu_e, la_e, h_p, gA, gF2out, gC, gF, gmu, gla);
    } else {
      constitutive_pre(k, C, F, mu_e, la_e, h_p, mat_p, o);
      constitutive_post(k, C, o);
      constitutive_bwd(k, C, F, o, gA, gF2out, gC, gF, gmu, gla);
    }
    float ox[3], ov[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      ox[d] = gx_in[d] + k.inv_dx * gfx[d];
      ov[d] = k.p_mass * S[d];
    }
    if (norm2) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ox[d] = nan_to_num(ox[d]);
        ov[d] = nan_to_num(ov[d]);
        gn2 += ox[d] * ox[d] + ov[d] * ov[d];
      }
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        gC.m[c] = nan_to_num(gC.m[c]);
        gF.m[c] = nan_to_num(gF.m[c]);
        gn2 += gC.m[c] * gC.m[c] + gF.m[c] * gF.m[c];
      }
    }
    {
      float og[PS_NCOMP];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        og[PS_X + d] = ox[d];
        og[PS_V + d] = ov[d];
      }
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        og[PS_C + c] = gC.m[c];
        og[PS_F + c] = gF.m[c];
      }
      store_comps<0, PS_NCOMP, PS_NQ>(gs, gp, og);
    }
  }
  // per-env scalars: one reduction per warp, then straight to the accumulators (no block barrier at the tail)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    gmu += __shfl_down_sync(0xffffffffu, gmu, off);
    gla += __shfl_down_sync(0xffffffffu, gla, off);
    gn2 += __shfl_down_sync(0xffffffffu, gn2, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&g_scal[env * GS_STRIDE + GS_MU], gmu);
    atomicAdd(&g_scal[env * GS_STRIDE + GS_LAMDA], gla);
    if (norm2) atomicAdd(&norm2[env * 2], gn2);
  }
  }  // tiles of this warp
}

void launch_p2g_bwd(const MpmConst& k, const float* ps_in, const float* svd_in, const float* mu_s,
                    const float* la_s, bool first_substep, const MpmWs& ws, cudaStream_t st, const float4* ggrid) {
  KScope ks_(KC_P2G_BWD, st);
  if (first_substep) cudaMemsetAsync(ws.norm2, 0, 4 * (size_t)k.B * 2, st);
  auto kern = k.liquid_fast ? k_p2g_bwd<true> : k_p2g_bwd<false>;
  kern<<<persistent_ctas(k, UD_BLOCK / 32, 4), UD_BLOCK, 0, st>>>(k, ps_in, svd_in, ggrid, ws.gs, mu_s, la_s, ws.mat_s, ws.h_s,
                                                                 ws.g_scal, first_substep ? ws.norm2 : nullptr);
}

// ------------------------------------------------------------------------------------------------
// norm_grad_state / norm_grad backward (mpm_simulator.py:389-408): nan_to_num every cotangent leaf,
// per-env global L2 norm over ALL leaves of the state, divide when the norm is >= 1.  The particle leaves are scrubbed
// and squared by the last k_p2g_bwd of the step (its `norm2` argument); the small leaves follow here.
// ------------------------------------------------------------------------------------------------
// one thread per env: small leaves (per-env scalars, primitive leaves, action).  Adds pass-through
// output cotangents, scrubs, accumulates the norms and stores the scrubbed values back.
__global__ void k_norm_small(MpmConst k, ud_mpm_state gout, float* __restrict__ g_scal,
                             float* __restrict__ g_prim_in, float* __restrict__ g_act,
                             float* __restrict__ norm2) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= k.B) return;
  float acc = 0.f, acc_a = 0.f;
  float* sc = g_scal + env * GS_STRIDE;
  sc[GS_FRICTION] = nan_to_num(sc[GS_FRICTION] + (gout.friction ? gout.friction[env] : 0.f));
  sc[GS_MU] = nan_to_num(sc[GS_MU] + (gout.mu ? gout.mu[env] : 0.f));
  sc[GS_LAMDA] = nan_to_num(sc[GS_LAMDA] + (gout.lamda ? gout.lamda[env] : 0.f));
  acc += sc[GS_FRICTION] * sc[GS_FRICTION] + sc[GS_MU] * sc[GS_MU] + sc[GS_LAMDA] * sc[GS_LAMDA];
  for (int q = 0; q < k.n_prim; ++q) {
    float* ps = sc + GS_PRIM + q * GS_PRIM_STRIDE;
    const ud_primitive& go = gout.prim[q];
    for (int j = 0; j < 3; ++j) {
      ps[j] = nan_to_num(ps[j] + (go.size ? go.size[env * 3 + j] : 0.f));
      acc += ps[j] * ps[j];
    }
    ps[3] = nan_to_num(ps[3] + (go.friction ? go.friction[env] : 0.f));
    acc += ps[3] * ps[3];
    float* pi = g_prim_in + ((size_t)env * k.n_prim + q) * 16;
    for (int j = 0; j < 13; ++j) {  // gpos0(3) grot0(4) gscale(6), pass-through already folded in
      pi[j] = nan_to_num(pi[j]);
      acc += pi[j] * pi[j];
    }
    float* ga = g_act + ((size_t)env * k.n_prim + q) * 6;
    for (int j = 0; j < 6; ++j) {
      ga[j] = nan_to_num(ga[j]);
      acc_a += ga[j] * ga[j];
    }
  }
  atomicAdd(&norm2[env * 2], acc);
  norm2[env * 2 + 1] = acc_a;
}

UD_DEV float norm_div(float g, float n2) {
  float nrm = sqrtf(n2);
  return nrm < 1.0f ? g : g / nrm;
}

__global__ void __launch_bounds__(256)
k_unsort_cot(MpmConst k, const float* __restrict__ gs, const int32_t* __restrict__ inv_perm,
             const float* __restrict__ norm2, float* __restrict__ gx, float* __restrict__ gv,
             float* __restrict__ gC, float* __restrict__ gF, float* __restrict__ gJ) {
  __shared__ float sm[256 / 32][32 * 9];
  UD_PARTICLE_INDEX(k, env, g);
  const int lane = threadIdx.x & 31;
  const int slot0 = slot_ - lane;
  if (slot0 >= k.n) return;                 // warp-uniform
  const int nlive = min(32, k.n - slot0);
  const size_t o0 = (size_t)env * k.n + slot0;   // one thread per ORIGINAL particle (coalesced AoS stores)
  const int sp = env * k.n_pad + (live_ ? inv_perm[g] : 0);
  float n2 = norm2[env * 2];
  float st[PS_NCOMP];
  load_comps<0, PS_NCOMP, PS_NQ>(gs, sp, st);
#pragma unroll
  for (int c = 0; c < PS_NCOMP; ++c) st[c] = norm_div(st[c], n2);
  float* wsm = sm[threadIdx.x >> 5];
  if (gx) warp_store_aos<3>(gx, o0, nlive, st + PS_X, wsm);
  if (gv) warp_store_aos<3>(gv, o0, nlive, st + PS_V, wsm);
  if (gC) warp_store_aos<9>(gC, o0, nlive, st + PS_C, wsm);
  if (gF) warp_store_aos<9>(gF, o0, nlive, st + PS_F, wsm);
  if (gJ && live_) gJ[g] = 0.f;  // the cotangent of J is dropped by substep_bwd_loss (:343-350)
}

__global__ void k_write_small(MpmConst k, ud_mpm_state gin, float* __restrict__ gaction,
                              const float* __restrict__ g_scal, const float* __restrict__ g_prim_in,
                              const float* __restrict__ g_act, const float* __restrict__ norm2) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= k.B) return;
  float n2 = norm2[env * 2], n2a = norm2[env * 2 + 1];
  const float* sc = g_scal + env * GS_STRIDE;
  if (gin.friction) gin.friction[env] = norm_div(sc[GS_FRICTION], n2);
  if (gin.mu) gin.mu[env] = norm_div(sc[GS_MU], n2);
  if (gin.lamda) gin.lamda[env] = norm_div(sc[GS_LAMDA], n2);
  const int S = k.S;
  for (int q = 0; q < k.n_prim; ++q) {
    const float* ps = sc + GS_PRIM + q * GS_PRIM_STRIDE;
    const float* pi = g_prim_in + ((size_t)env * k.n_prim + q) * 16;
    const ud_primitive& gi = gin.prim[q];
    if (gi.size)
      for (int j = 0; j < 3; ++j) gi.size[env * 3 + j] = norm_div(ps[j], n2);
    if (gi.friction) gi.friction[env] = norm_div(ps[3], n2);
    // only row 0 of the input tables reaches the output (rows >= 1 and v, w, action_buffer are
    // overwritten by FK / set_action)
    for (int f = 0; f < S; ++f) {
      for (int j = 0; j < 3; ++j) {
        if (gi.position) gi.position[((size_t)env * S + f) * 3 + j] = f == 0 ? norm_div(pi[j], n2) : 0.f;
        if (gi.v) gi.v[((size_t)env * S + f) * 3 + j] = 0.f;
        if (gi.w) gi.w[((size_t)env * S + f) * 3 + j] = 0.f;
      }
      for (int j = 0; j < 4; ++j)
        if (gi.rotation) gi.rotation[((size_t)env * S + f) * 4 + j] = f == 0 ? norm_div(pi[3 + j], n2) : 0.f;
    }
    for (int j = 0; j < 6; ++j) {
      if (gi.action_scale) gi.action_scale[env * 6 + j] = norm_div(pi[7 + j], n2);
      if (gi.action_buffer) gi.action_buffer[env * 6 + j] = 0.f;
      if (gaction) gaction[(size_t)env * 6 * k.n_prim + 6 * q + j] =
          norm_div(g_act[((size_t)env * k.n_prim + q) * 6 + j], n2a);
    }
  }
}

void launch_finish_bwd(const MpmConst& k, const ud_mpm_state* in, const ud_mpm_state* gout,
                       ud_mpm_state* gin, const float* action, float* gaction, const MpmWs& ws,
                       cudaStream_t st) {
  KScope ks_(KC_FINISH_BWD, st, 5);
  (void)in;
  (void)action;
  // the particle part of the norm (and the nan_to_num of the particle cotangents) was done by the last k_p2g_bwd
  k_norm_small<<<cdiv(k.B, 64), 64, 0, st>>>(k, *gout, ws.g_scal, ws.g_prim_in, ws.g_act, ws.norm2);
  k_unsort_cot<<<pgrid(k, 256), 256, 0, st>>>(k, ws.gs, ws.inv_perm, ws.norm2, gin->x, gin->v, gin->C, gin->F,
                                              gin->J);
  k_write_small<<<cdiv(k.B, 64), 64, 0, st>>>(k, *gin, gaction, ws.g_scal, ws.g_prim_in, ws.g_act, ws.norm2);
}

}  // namespace ud
