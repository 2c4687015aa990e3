// Particle-side kernels of the MLS-MPM step: per-frame binning, P2G scatter, G2P gather and their
// adjoints.  sm_100a.  Reference: DaXBench/daxbench/core/engine/mpm_simulator.py:178-330.
#include "mpm_internal.h"

namespace ud {

#define UD_BLOCK 128

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

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
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= k.N) return;
  int env = g / k.n;
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

// one CTA per env: exclusive scan of NK counts in place (count -> start), start[NK] = n
__global__ void k_scan(MpmConst k, int32_t* __restrict__ cell_start) {
  __shared__ int part[1024];
  int32_t* c = cell_start + (size_t)blockIdx.x * (k.NK + 1);
  int per = (k.NK + blockDim.x - 1) / blockDim.x;
  int lo = threadIdx.x * per, hi = min(lo + per, k.NK);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += c[i];
  part[threadIdx.x] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over blockDim partials
  for (int off = 1; off < blockDim.x; off <<= 1) {
    int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - s;
  for (int i = lo; i < hi; ++i) {
    int t = c[i];
    c[i] = run;
    run += t;
  }
  if (threadIdx.x == blockDim.x - 1) c[k.NK] = part[threadIdx.x];
}

__global__ void k_place(MpmConst k, const int32_t* __restrict__ keys, const int32_t* __restrict__ cell_start,
                        int32_t* __restrict__ cursor, int32_t* __restrict__ tmp_idx) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= k.N) return;
  int env = g / k.n;
  int key = keys[g];
  int slot = cell_start[(size_t)env * (k.NK + 1) + key] + atomicAdd(&cursor[(size_t)env * k.NK + key], 1);
  tmp_idx[(size_t)env * k.n + slot] = g - env * k.n;
}

__global__ void k_rank(MpmConst k, const int32_t* __restrict__ keys, const int32_t* __restrict__ cell_start,
                       const int32_t* __restrict__ tmp_idx, int32_t* __restrict__ perm) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= k.N) return;
  int env = g / k.n;
  int p = tmp_idx[g];
  int key = keys[env * k.n + p];
  const int32_t* cs = cell_start + (size_t)env * (k.NK + 1);
  int lo = cs[key], hi = cs[key + 1];
  const int32_t* seg = tmp_idx + (size_t)env * k.n;
  int rank = 0;
  for (int j = lo; j < hi; ++j) rank += seg[j] < p;
  perm[(size_t)env * k.n + lo + rank] = p;
}

void launch_sort(const MpmConst& k, const float* x_aos, const MpmWs& ws, int32_t* out_base, cudaStream_t st) {
  cudaMemsetAsync(ws.cell_start, 0, sizeof(int32_t) * (size_t)k.B * (k.NK + 1), st);
  cudaMemsetAsync(ws.cursor, 0, sizeof(int32_t) * (size_t)k.B * k.NK, st);
  k_keys<<<cdiv(k.N, 256), 256, 0, st>>>(k, x_aos, ws.keys, ws.cell_start, out_base);
  k_scan<<<k.B, 1024, 0, st>>>(k, ws.cell_start);
  k_place<<<cdiv(k.N, 256), 256, 0, st>>>(k, ws.keys, ws.cell_start, ws.cursor, ws.tmp_idx);
  k_rank<<<cdiv(k.N, 256), 256, 0, st>>>(k, ws.keys, ws.cell_start, ws.tmp_idx, ws.perm);
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
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= k.N) return;
  int env = g / k.n;
  int p = perm[g];
  size_t o = (size_t)env * k.n + p;
  const size_t N = k.N;
#pragma unroll
  for (int c = 0; c < 3; ++c) ps[(PS_X + c) * N + g] = nan_to_num(x[3 * o + c]);
#pragma unroll
  for (int c = 0; c < 3; ++c) ps[(PS_V + c) * N + g] = nan_to_num(v[3 * o + c]);
#pragma unroll
  for (int c = 0; c < 9; ++c) ps[(PS_C + c) * N + g] = nan_to_num(C[9 * o + c]);
#pragma unroll
  for (int c = 0; c < 9; ++c) ps[(PS_F + c) * N + g] = nan_to_num(F[9 * o + c]);
  mat_s[g] = material[p];
  h_s[g] = h[p];
}

void launch_gather_state(const MpmConst& k, const ud_mpm_state* in, const int32_t* material, const float* h,
                         const MpmWs& ws, float* ps_slot, cudaStream_t st) {
  k_gather_state<<<cdiv(k.N, 256), 256, 0, st>>>(k, in->x, in->v, in->C, in->F, material, h, ws.perm, ps_slot,
                                                  ws.mat_s, ws.h_s);
}

// ------------------------------------------------------------------------------------------------
// P2G: F update + SVD + plasticity + stress (mpm_simulator.py:238-268), then the 27-node scatter of
// {momentum, mass} (p2g_micro, :178-194) as ONE 16-byte vector reduction per node
// (red.global.add.v4.f32 -> REDG.E.ADD.F32x4).  Out-of-range nodes are dropped (JAX scatter rule).
// ------------------------------------------------------------------------------------------------
UD_DEV void load_particle(const float* __restrict__ ps, size_t N, int g, float x[3], float v[3], Mat3& C,
                          Mat3& F) {
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = ps[(PS_X + c) * N + g];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = ps[(PS_V + c) * N + g];
#pragma unroll
  for (int c = 0; c < 9; ++c) C.m[c] = ps[(PS_C + c) * N + g];
#pragma unroll
  for (int c = 0; c < 9; ++c) F.m[c] = ps[(PS_F + c) * N + g];
}

__global__ void __launch_bounds__(UD_BLOCK)
k_p2g(MpmConst k, const float* __restrict__ ps_in, float* __restrict__ ps_out, float4* __restrict__ grid,
      const float* __restrict__ mu_s, const float* __restrict__ la_s, const int32_t* __restrict__ mat_s,
      const float* __restrict__ h_s) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= k.N) return;
  const size_t N = k.N;
  int env = g / k.n;
  float x[3], v[3];
  Mat3 C, F;
  load_particle(ps_in, N, g, x, v, C, F);
  Stencil st;
  make_stencil(x, k.inv_dx, st);
  Consti o;
  constitutive_fwd(k, C, F, mu_s[env], la_s[env], h_s[g], mat_s[g], o);
#pragma unroll
  for (int c = 0; c < 9; ++c) ps_out[(PS_F + c) * N + g] = o.F2.m[c];
  float4* genv = grid + (size_t)env * k.G;
  float mv[3] = {k.p_mass * v[0], k.p_mass * v[1], k.p_mass * v[2]};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int ix = idx_scatter(st.base[0] + a, k.rx);
    float dx0 = ((float)a - st.fx[0]) * k.dx;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      int iy = idx_scatter(st.base[1] + b, k.ry);
      float dx1 = ((float)b - st.fx[1]) * k.dx;
      float wab = st.w[a][0] * st.w[b][1];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        int iz = idx_scatter(st.base[2] + c, k.rz);
        if ((ix | iy | iz) < 0) continue;
        float dx2 = ((float)c - st.fx[2]) * k.dx;
        float wt = wab * st.w[c][2];
        float4 val;
        val.x = wt * (mv[0] + (o.affine(0, 0) * dx0 + o.affine(0, 1) * dx1 + o.affine(0, 2) * dx2));
        val.y = wt * (mv[1] + (o.affine(1, 0) * dx0 + o.affine(1, 1) * dx1 + o.affine(1, 2) * dx2));
        val.z = wt * (mv[2] + (o.affine(2, 0) * dx0 + o.affine(2, 1) * dx1 + o.affine(2, 2) * dx2));
        val.w = wt * k.p_mass;
        atomicAdd(&genv[(ix * k.ry + iy) * k.rz + iz], val);
      }
    }
  }
}

void launch_p2g(const MpmConst& k, const float* ps_in, float* ps_out, float4* grid, const float* mu_s,
                const float* la_s, const MpmWs& ws, cudaStream_t st) {
  k_p2g<<<cdiv(k.N, UD_BLOCK), UD_BLOCK, 0, st>>>(k, ps_in, ps_out, grid, mu_s, la_s, ws.mat_s, ws.h_s);
}

// ------------------------------------------------------------------------------------------------
// G2P (g2p_micro, :196-221) + advection (:326).  Out-of-range nodes clamp (JAX gather rule).
// Rows of C' of original particles 0..2 are kept for the J update quirk (:327).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UD_BLOCK)
k_g2p(MpmConst k, const float* __restrict__ ps_in, float* __restrict__ ps_out, const float4* __restrict__ grid,
      const int32_t* __restrict__ perm, float* __restrict__ jrows, int substep) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= k.N) return;
  const size_t N = k.N;
  int env = g / k.n;
  float x[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = ps_in[(PS_X + c) * N + g];
  Stencil st;
  make_stencil(x, k.inv_dx, st);
  const float4* genv = grid + (size_t)env * k.G;
  float nv[3] = {0.f, 0.f, 0.f};
  Mat3 nC = mat_zero();
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int ix = idx_gather(st.base[0] + a, k.rx);
    float d0 = (float)a - st.fx[0];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      int iy = idx_gather(st.base[1] + b, k.ry);
      float d1 = (float)b - st.fx[1];
      float wab = st.w[a][0] * st.w[b][1];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        int iz = idx_gather(st.base[2] + c, k.rz);
        float d2 = (float)c - st.fx[2];
        float wt = wab * st.w[c][2];
        float4 gv = __ldg(&genv[(ix * k.ry + iy) * k.rz + iz]);
        nv[0] += wt * gv.x;
        nv[1] += wt * gv.y;
        nv[2] += wt * gv.z;
        float w4 = 4.f * wt;
        float gvv[3] = {gv.x, gv.y, gv.z};
        float dd[3] = {d0, d1, d2};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) nC(i, j) += w4 * (gvv[i] * dd[j]) * k.inv_dx;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) ps_out[(PS_X + c) * N + g] = x[c] + k.dt * nv[c];
#pragma unroll
  for (int c = 0; c < 3; ++c) ps_out[(PS_V + c) * N + g] = nv[c];
#pragma unroll
  for (int c = 0; c < 9; ++c) ps_out[(PS_C + c) * N + g] = nC.m[c];
  int p = perm[g];
  if (p < 3) {
    float* jr = jrows + ((size_t)env * k.S + substep) * 9 + p * 3;
    jr[0] = nC(p, 0);
    jr[1] = nC(p, 1);
    jr[2] = nC(p, 2);
  }
}

void launch_g2p(const MpmConst& k, const float* ps_in, float* ps_out, const float4* grid, int substep,
                const MpmWs& ws, cudaStream_t st) {
  k_g2p<<<cdiv(k.N, UD_BLOCK), UD_BLOCK, 0, st>>>(k, ps_in, ps_out, grid, ws.perm, ws.jrows, substep);
}

// sorted SoA -> AoS outputs; J' = J * prod_f (1 + dt * trace-quirk_f), sequentially as the reference
__global__ void k_unsort_state(MpmConst k, const float* __restrict__ ps, const float* __restrict__ J_in,
                               const int32_t* __restrict__ perm, const float* __restrict__ jrows,
                               float* __restrict__ x, float* __restrict__ v, float* __restrict__ C,
                               float* __restrict__ F, float* __restrict__ J) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= k.N) return;
  const size_t N = k.N;
  int env = g / k.n;
  int p = perm[g];
  size_t o = (size_t)env * k.n + p;
#pragma unroll
  for (int c = 0; c < 3; ++c) x[3 * o + c] = ps[(PS_X + c) * N + g];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[3 * o + c] = ps[(PS_V + c) * N + g];
#pragma unroll
  for (int c = 0; c < 9; ++c) C[9 * o + c] = ps[(PS_C + c) * N + g];
#pragma unroll
  for (int c = 0; c < 9; ++c) F[9 * o + c] = ps[(PS_F + c) * N + g];
  float j = nan_to_num(J_in[o]);
  int nr = min(3, k.n);
  for (int f = 0; f < k.S; ++f) {
    const float* jr = jrows + ((size_t)env * k.S + f) * 9;
    float t[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < nr; ++i) {
      t[0] += jr[3 * i];
      t[1] += jr[3 * i + 1];
      t[2] += jr[3 * i + 2];
    }
    float tr = (t[0] + t[1]) + t[2];
    j = j * (1.f + k.dt * tr);
  }
  J[o] = j;
}

void launch_unsort_state(const MpmConst& k, const float* ps_slot, const float* J_in, const MpmWs& ws,
                         ud_mpm_state* out, cudaStream_t st) {
  k_unsort_state<<<cdiv(k.N, 256), 256, 0, st>>>(k, ps_slot, J_in, ws.perm, ws.jrows, out->x, out->v, out->C,
                                                  out->F, out->J);
}

}  // namespace ud
