// Grid-side kernels of the MLS-MPM step: primitive forward kinematics, the fused grid update
// (normalise + gravity + colliders + ground friction + boundary) and their adjoints.
// sm_100a; compile this file with --fmad=false (see mpm_cell.cuh).
// Reference: mpm_simulator.py:277-313,365-373,413-429; primitives.py:73-92,185-229.
#include "mpm_cell.cuh"
#include "mpm_internal.h"

namespace ud {

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// primitives.py:73-81 (terms = outer(r, q)), normalised
template <class T> UD_DEV void qmul(const T q[4], const T r[4], T out[4]) {
  T w = r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3];
  T x = r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2];
  T y = r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1];
  T z = r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0];
  T nrm = s_max_c(s_sqrt(w * w + x * x + y * y + z * z), 1e-12f);
  out[0] = w / nrm;
  out[1] = x / nrm;
  out[2] = y / nrm;
  out[3] = z / nrm;
}
// primitives.py:84-92
template <class T> UD_DEV void w2quat(const T aa[3], T out[4]) {
  T w = s_sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]) + 1e-12f;
  T sn = s_sin(w / 2.f);
  out[0] = s_cos(w / 2.f);
  out[1] = (aa[0] / w) * sn;
  out[2] = (aa[1] / w) * sn;
  out[3] = (aa[2] / w) * sn;
}

UD_DEV float clipf(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }
// jnp.clip = minimum(hi, maximum(lo, a)); lax.max/min split the cotangent evenly at a tie
UD_DEV float clip_grad(float a, float lo, float hi) {
  return (a > lo && a < hi) ? 1.f : ((a == lo || a == hi) ? 0.5f : 0.f);
}

// ------------------------------------------------------------------------------------------------
// step prologue for the primitives (mpm_simulator.py:419-423 + the FK of every substep, which only
// depends on the action): action clip, set_action, rows 1..S-1 of position/rotation, copy_frame.
// One thread per (env, primitive); S is 16..133.
// Row f+1 == S is never written and reads of it clamp to S-1 (JAX scatter drop / gather clamp).
// ------------------------------------------------------------------------------------------------
__global__ void k_fk_fwd(MpmConst k, ud_mpm_state in, const float* __restrict__ action, ud_mpm_state out,
                         float* __restrict__ fk_pos, float* __restrict__ fk_rot, float* __restrict__ fk_vw,
                         float* __restrict__ fk_act, int write_out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k.B * k.n_prim) return;
  int env = t / k.n_prim, q = t % k.n_prim;
  const ud_primitive& pi = in.prim[q];
  const ud_primitive& po = out.prim[q];
  const int S = k.S;
  float a[6], vw[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    a[j] = clipf(action[(size_t)env * 6 * k.n_prim + 6 * q + j], -1.f, 1.f);
    vw[j] = a[j] * pi.action_scale[env * 6 + j] / (float)S;
    fk_act[(size_t)t * 6 + j] = a[j];
    fk_vw[(size_t)t * 6 + j] = vw[j];
  }
  float* tp = fk_pos + (size_t)t * (S + 1) * 3;
  float* tr = fk_rot + (size_t)t * (S + 1) * 4;
  float pos[3], rot[4], dq[4];
#pragma unroll
  for (int j = 0; j < 3; ++j) pos[j] = pi.position[((size_t)env * S) * 3 + j];
#pragma unroll
  for (int j = 0; j < 4; ++j) rot[j] = pi.rotation[((size_t)env * S) * 4 + j];
  w2quat(vw + 3, dq);
  for (int f = 0; f < S; ++f) {
#pragma unroll
    for (int j = 0; j < 3; ++j) tp[f * 3 + j] = pos[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) tr[f * 4 + j] = rot[j];
    if (f + 1 < S) {
#pragma unroll
      for (int j = 0; j < 3; ++j) pos[j] = clipf(pos[j] + vw[j], -2.f, 2.f);
      float nr[4];
      qmul(dq, rot, nr);
#pragma unroll
      for (int j = 0; j < 4; ++j) rot[j] = nr[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) tp[S * 3 + j] = pos[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) tr[S * 4 + j] = rot[j];
  if (!write_out) return;
  // output leaves: rows 1..S-1 as computed, row 0 = copy_frame(S -> 0) = row S-1
  for (int f = 0; f < S; ++f) {
    int src = f == 0 ? S - 1 : f;
#pragma unroll
    for (int j = 0; j < 3; ++j) po.position[((size_t)env * S + f) * 3 + j] = tp[src * 3 + j];
#pragma unroll
    for (int j = 0; j < 4; ++j) po.rotation[((size_t)env * S + f) * 4 + j] = tr[src * 4 + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      po.v[((size_t)env * S + f) * 3 + j] = vw[j];
      po.w[((size_t)env * S + f) * 3 + j] = vw[3 + j];
    }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    po.action_buffer[env * 6 + j] = a[j];
    po.action_scale[env * 6 + j] = pi.action_scale[env * 6 + j];
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) po.size[env * 3 + j] = pi.size[env * 3 + j];
  po.friction[env] = pi.friction[env];
  po.softness[env] = pi.softness[env];
}

void launch_fk_fwd(const MpmConst& k, const ud_mpm_state* in, const float* action, ud_mpm_state* out,
                   const MpmWs& ws, cudaStream_t st) {
  int n = k.B * k.n_prim;
  if (n == 0) return;
  KScope ks_(KC_FK, st);
  k_fk_fwd<<<cdiv(n, 64), 64, 0, st>>>(k, *in, action, out ? *out : *in, ws.fk_pos, ws.fk_rot, ws.fk_vw,
                                       ws.fk_act, out != nullptr);
}

// ------------------------------------------------------------------------------------------------
// Fused grid update, one thread per cell, dense over B*G.  Cells that received no mass are skipped:
// nothing gathers them with a non-zero weight (the reference computes gravity/collider values there
// that are never read).
// ------------------------------------------------------------------------------------------------
template <class T>
UD_DEV void load_prim_f(const MpmConst& k, const ud_mpm_state& in, const float* fk_pos, const float* fk_rot,
                        const float* fk_vw, int env, int q, int f, PrimIn<float>& pr) {
  size_t t = (size_t)env * k.n_prim + q;
  const float* tp = fk_pos + t * (k.S + 1) * 3;
  const float* tr = fk_rot + t * (k.S + 1) * 4;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    pr.pos_f[j] = tp[f * 3 + j];
    pr.pos_f1[j] = tp[(f + 1) * 3 + j];
    pr.size[j] = in.prim[q].size[env * 3 + j];
    pr.v_f[j] = fk_vw[t * 6 + j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    pr.rot_f[j] = tr[f * 4 + j];
    pr.rot_f1[j] = tr[(f + 1) * 4 + j];
  }
  pr.friction = in.prim[q].friction[env];
  pr.softness = in.prim[q].softness[env];
}

// Compacts the block marks P2G left (ws.blk_flag) into a list and consumes them: one warp per 32 blocks, one
// warp-aggregated append.  ~3-6 % of the blocks are marked in the shipped scenes.
__global__ void __launch_bounds__(128)
k_blk_compact(int total, int32_t* __restrict__ blk_flag, int32_t* __restrict__ list, int32_t* __restrict__ count, int stamp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool marked = i < total && blk_flag[i] == BLK_MARK_CTA;
  if (marked) blk_flag[i] = stamp;   // same state the self-listing warp kernels leave: flag == substep + 1 <=> listed
  if (i == 0) count[1] = 1;           // the CTA-staged kernels do not look for stencils that leave the grid: shell job on
  const unsigned m = __ballot_sync(0xffffffffu, marked);
  if (!m) return;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (marked) list[base + __popc(m & ((1u << lane) - 1u))] = i;
}

// update of one EMPTY boundary cell (raw value g, mass 0)
__device__ __forceinline__ void
empty_shell_cell(const MpmConst& k, int ci, int cj, int ck, size_t idx, int env, const float4& g, float4* grid_out, int f,
                 const ud_mpm_state& in, const float* __restrict__ fk_pos, const float* __restrict__ fk_rot,
                 const float* __restrict__ fk_vw, int32_t* __restrict__ act_list, int32_t* __restrict__ act_count,
                 float4* __restrict__ act_raw) {
  const float gpos[3] = {(float)ci * k.dx, (float)cj * k.dx, (float)ck * k.dx};
  float p[3] = {g.x, g.y, g.z}, v[3];
  bool any_active = false;
  auto prim_of = [&](int q, PrimIn<float>& pr) {
    load_prim_f<float>(k, in, fk_pos, fk_rot, fk_vw, env, q, f, pr);
    const bool a = prim_active(k, gpos, pr);  // without influence the primitive changes v by < 1e-12 relative: skipped
    any_active |= a;
    return a;
  };
  cell_update<float>(k, ci, cj, ck, p, g.w, in.friction[env], prim_of, v);
  grid_out[idx] = make_float4(v[0], v[1], v[2], g.w);
  // an empty shell cell under a primitive's influence can carry a cotangent to that primitive: list it for k_grid_bwd
  if (act_list && any_active) {
    const int at = atomicAdd(act_count, 1);
    act_list[at] = (int32_t)idx;
    act_raw[at] = g;
  }
}

// Grid update over the listed 4x4x4 blocks: persistent warps, one block (64 cells, 2 per lane) per iteration.
__device__ __forceinline__ void
grid_fwd_body(const MpmConst& k, float4* grid_in, float4* grid_out, long long* __restrict__ grid_fix, int f,
              const ud_mpm_state& in, const float* __restrict__ fk_pos, const float* __restrict__ fk_rot,
              const float* __restrict__ fk_vw, const int32_t* __restrict__ blk_list, const int32_t* __restrict__ blk_count,
              int32_t* __restrict__ act_list, int32_t* __restrict__ act_count, float4* __restrict__ act_raw, int vblock,
              int nblocks) {
  // (Running the update over the COMPACTED cells of a block -- ballot, __fns, shuffles of the raw values -- so that a
  // block with fewer than 32 cells to update costs one dependent chain instead of two was measured and not kept: grid
  // 14.8 -> 15.5 us per launch in the plasticine scene, 66.0 -> 66.6 us in pour_water; profiles/r02_run39.sh.)
  const int nblk = k.nbx * k.nby * k.nbz;
  const int lane = threadIdx.x & 31;
  const int warp0 = (int)((vblock * (size_t)blockDim.x + threadIdx.x) >> 5), nwarps = (nblocks * blockDim.x) >> 5;
  const int count = *blk_count;
  for (int li = warp0; li < count; li += nwarps) {
  const int w = blk_list[li];
  const int env = w / nblk, bid = w - env * nblk;
  const int bz = bid % k.nbz, by = (bid / k.nbz) % k.nby, bx = bid / (k.nbz * k.nby);
#pragma unroll 1
  for (int it = 0; it < 2; ++it) {
    const int lc = lane + 32 * it;
    const int ci = bx * 4 + (lc >> 4), cj = by * 4 + ((lc >> 2) & 3), ck = bz * 4 + (lc & 3);
    const bool in_range = ci < k.rx && cj < k.ry && ck < k.rz;
    const size_t idx = (size_t)env * k.G + (size_t)(ci * k.ry + cj) * k.rz + ck;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_range) {
      if (grid_fix) {  // deterministic P2G: fixed-point accumulators -> float {p, m}; consumed (re-zeroed) here
        longlong2* fx = reinterpret_cast<longlong2*>(grid_fix) + 2 * idx;
        const longlong2 a = fx[0], b = fx[1];
        fx[0] = make_longlong2(0, 0);
        fx[1] = make_longlong2(0, 0);
        g = make_float4((float)((double)a.x * FIX_INV), (float)((double)a.y * FIX_INV), (float)((double)b.x * FIX_INV),
                        (float)((double)b.y * FIX_INV));
      } else {
        g = grid_in[idx];
      }
    }
    const bool has_mass = in_range && g.w > 0.f;
    if (act_list) {  // recompute pass of the adjoint: list the cells with mass for k_grid_bwd (warp-aggregated append)
      const unsigned m = __ballot_sync(0xffffffffu, has_mass);
      if (m) {
        int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(act_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (has_mass) {   // the raw {p, m} travels with the list: the raw grids themselves are re-used every other substep
          const int at = base + __popc(m & ((1u << lane) - 1u));
          act_list[at] = (int32_t)idx;
          act_raw[at] = g;
        }
      }
    }
    const bool on_shell = ci == 0 || cj == 0 || ck == 0 || ci == k.rx - 1 || cj == k.ry - 1 || ck == k.rz - 1;
    const float gpos[3] = {(float)ci * k.dx, (float)cj * k.dx, (float)ck * k.dx};
    // Scenes with several primitives (pour_water: two bowls, each out of reach of the other's liquid): a primitive that
    // no cell of this half-block is within reach of (influence below 1e-12, i.e. 1 - influence rounds to 1; the same
    // test the empty cells and the adjoint use) is skipped for the whole warp -- a warp-uniform decision, no divergence.
    // With ONE primitive the test does not pay (see below) and every cell with mass evaluates it, as in the reference.
    unsigned reach = 0xffffffffu;
    if (k.n_prim >= 2) {   // kernel-uniform; all 32 lanes are here (branch-free per lane: votes with the full mask)
      reach = 0u;
      const bool work = in_range && (has_mass || on_shell);
#pragma unroll 1
      for (int q = 0; q < k.n_prim; ++q) {
        PrimIn<float> pr;
        load_prim_f<float>(k, in, fk_pos, fk_rot, fk_vw, env, q, f, pr);
        const bool a = prim_active(k, gpos, pr) && work;
        if (__any_sync(0xffffffffu, a)) reach |= 1u << q;
      }
    }
    if (!in_range) continue;
    if (!has_mass && !on_shell) {  // empty interior cell: never gathered with a non-zero weight
      if (grid_out != grid_in || grid_fix) grid_out[idx] = g;
      continue;
    }
    // ONE inlined copy of the cell update for both kinds of cell (the kernel's code size is what bounds it, see
    // k_grid_fwd).  A cell with mass evaluates every primitive, as in the reference.  (Skipping primitives whose
    // influence is below 1e-12 was measured: in the plasticine scene half of the cells with mass are within reach of
    // the pusher, nearly every 4x4x4 block has an active lane, and the extra test made the kernel 25 % longer.)
    // Every cell of a listed block is this job's, its EMPTY boundary cells included (the shell job skips listed
    // blocks): those skip primitives without influence and are listed for k_grid_bwd when one acts on them.
    float p[3] = {g.x, g.y, g.z}, v[3];
    bool any_active = false;
    auto prim_of = [&](int q, PrimIn<float>& pr) {
      if (!((reach >> q) & 1u)) return false;   // warp-uniform
      load_prim_f<float>(k, in, fk_pos, fk_rot, fk_vw, env, q, f, pr);
      if (has_mass) return true;
      const bool a = prim_active(k, gpos, pr);
      any_active |= a;
      return a;
    };
    cell_update<float>(k, ci, cj, ck, p, g.w, in.friction[env], prim_of, v);
    grid_out[idx] = make_float4(v[0], v[1], v[2], g.w);
    if (!has_mass && act_list && any_active) {
      const int at = atomicAdd(act_count, 1);
      act_list[at] = (int32_t)idx;
      act_raw[at] = g;
    }
  }
  }  // listed blocks of this warp
}

// Face cell number t of the grid's outermost layer (z faces, then y faces, then x faces; edges appear twice).
__device__ __forceinline__ bool shell_cell(const MpmConst& k, int t, int& ci, int& cj, int& ck) {
  const int nxy = k.rx * k.ry, nxz = k.rx * k.rz, nyz = k.ry * k.rz;
  if (t < 2 * nxy) {
    ck = t >= nxy ? k.rz - 1 : 0;
    t -= t >= nxy ? nxy : 0;
    ci = t / k.ry;
    cj = t % k.ry;
  } else if (t < 2 * (nxy + nxz)) {
    t -= 2 * nxy;
    cj = t >= nxz ? k.ry - 1 : 0;
    t -= t >= nxz ? nxz : 0;
    ci = t / k.rz;
    ck = t % k.rz;
  } else if (t < 2 * (nxy + nxz + nyz)) {
    t -= 2 * (nxy + nxz);
    ci = t >= nyz ? k.rx - 1 : 0;
    t -= t >= nyz ? nyz : 0;
    cj = t / k.rz;
    ck = t % k.rz;
  } else {
    return false;
  }
  return true;
}

// Between substeps, instead of a memset of the whole grid (the forward's raw grids and the adjoint's cotangent grids
// are both used alternately, and the grid-side launch of a substep re-zeroes the slot the next one scatters into):
// re-zero (a) the face cells written outside the listed blocks -- CTAs [0, shell_ctas), only in substeps whose shell
// flag is up -- and (b) the blocks that substep's P2G touched -- the remaining CTAs, persistent warps over the list.
__device__ __forceinline__ void
grid_clear_body(const MpmConst& k, float4* __restrict__ grid, const int32_t* __restrict__ blk_list,
                const int32_t* __restrict__ blk_count, int shell_ctas_per_env, int vblock, int nblocks) {
  const int shell_ctas = shell_ctas_per_env * k.B;
  if (vblock < shell_ctas) {
    if (blk_count[1] == 0) return;   // no stencil left the grid in that substep: no face cell outside the listed blocks was written
    const int env = vblock / shell_ctas_per_env;
    int ci, cj, ck;
    if (!shell_cell(k, (vblock - env * shell_ctas_per_env) * blockDim.x + threadIdx.x, ci, cj, ck)) return;
    grid[(size_t)env * k.G + (size_t)(ci * k.ry + cj) * k.rz + ck] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int nblk = k.nbx * k.nby * k.nbz;
  const int lane = threadIdx.x & 31;
  const int warp0 = (int)((((size_t)vblock - shell_ctas) * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)(((nblocks - shell_ctas) * blockDim.x) >> 5);
  const int count = *blk_count;
  for (int li = warp0; li < count; li += nwarps) {
    const int w = blk_list[li];
    const int env = w / nblk, bid = w - env * nblk;
    const int bz = bid % k.nbz, by = (bid / k.nbz) % k.nby, bx = bid / (k.nbz * k.nby);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int lc = lane + 32 * it;
      const int ci = bx * 4 + (lc >> 4), cj = by * 4 + ((lc >> 2) & 3), ck = bz * 4 + (lc & 3);
      if (ci < k.rx && cj < k.ry && ck < k.rz)
        grid[(size_t)env * k.G + (size_t)(ci * k.ry + cj) * k.rz + ck] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Empty cells of the outermost layer: the reference updates EVERY cell (empty ones get dt*gravity, colliders,
// friction, walls), and a particle that left the grid gathers exactly these cells through clamped indices
// (SURVEY 8c).  Cells of a block P2G listed this substep are updated by the block job (which calls empty_shell_cell
// for them); the shell job, one thread per face cell, covers the blocks nobody scattered into.  Ownership is decided
// by the block flag alone, never by re-reading a grid value another job may be writing.
__device__ __forceinline__ void
grid_shell_body(const MpmConst& k, const float4* grid_raw, float4* grid_out, int f, const ud_mpm_state& in,
                const float* __restrict__ fk_pos, const float* __restrict__ fk_rot, const float* __restrict__ fk_vw,
                int32_t* __restrict__ act_list, int32_t* __restrict__ act_count, float4* __restrict__ act_raw,
                const int32_t* __restrict__ blk_flag, int env, int vblock) {
  const int t0 = vblock * blockDim.x + threadIdx.x;
  const int nxy = k.rx * k.ry, nxz = k.rx * k.rz;
  int ci, cj, ck;
  if (!shell_cell(k, t0, ci, cj, ck)) return;
  // edge and corner cells belong to two or three faces: ONE face owns each (z faces first, then y, then x)
  const bool owner = (ck == 0 || ck == k.rz - 1) ? (t0 < 2 * nxy)
                     : ((cj == 0 || cj == k.ry - 1) ? (t0 >= 2 * nxy && t0 < 2 * (nxy + nxz)) : true);
  if (!owner) return;
  if (blk_flag[(size_t)env * (k.nbx * k.nby * k.nbz) + ((ci >> 2) * k.nby + (cj >> 2)) * k.nbz + (ck >> 2)] == f + 1)
    return;  // a listed block: the block job's cell
  const size_t idx = (size_t)env * k.G + (size_t)(ci * k.ry + cj) * k.rz + ck;
  // nobody scattered into this block: the cell is empty, and grid_raw holds what the call's memset / the re-zeroing
  // left there (zero)
  empty_shell_cell(k, ci, cj, ck, idx, env, grid_raw[idx], grid_out, f, in, fk_pos, fk_rot, fk_vw, act_list, act_count, act_raw);
}

// ONE launch per substep for the whole grid side: CTAs [0, NF) update the listed 4x4x4 blocks (persistent warps),
// the next `shell` CTAs update the boundary cells of the blocks that are NOT listed, and -- in the in-place forward --
// the remaining CTAs re-zero the OTHER grid (the one the next substep scatters into) from the block list of the
// previous substep.  The three jobs touch disjoint cells.
constexpr int GRID_FWD_CTAS = 148 * 8;
struct GridClearArgs {
  float4* grid;             // null: nothing to clear
  const int32_t* blk_list;
  const int32_t* blk_count;
  int ctas;                 // shell_ctas_per_env * B + persistent CTAs
  int shell_ctas_per_env;
};
// KIND / PC: the collider's SDF (UD_SDF_*) and position control as COMPILE-TIME constants.  Both are call parameters, and
// with both variants inlined at each of the 7 SDF evaluations per primitive the kernel was 9 768 (k_grid_bwd: 13 088)
// instructions = 156 KB (209 KB) of code that every warp walks through exactly once: ncu showed "no instruction"
// (instruction-cache misses) as the top stall of both kernels, 4.9 cycles per issue in k_grid_bwd.
template <int KIND, int PC>
__global__ void __launch_bounds__(128)
k_grid_fwd(MpmConst k_, float4* grid_in, float4* grid_out, long long* __restrict__ grid_fix, int f,
           ud_mpm_state in, const float* __restrict__ fk_pos, const float* __restrict__ fk_rot,
           const float* __restrict__ fk_vw, const int32_t* __restrict__ blk_list, const int32_t* __restrict__ blk_count,
           int32_t* __restrict__ act_list, int32_t* __restrict__ act_count, float4* __restrict__ act_raw,
           const int32_t* __restrict__ blk_flag, int shell_ctas_per_env, GridClearArgs clr) {
  MpmConst k = k_;
  k.sdf_kind = KIND;
  k.pos_control = PC;
  int vb = blockIdx.x;
  if (vb < GRID_FWD_CTAS) {
    grid_fwd_body(k, grid_in, grid_out, grid_fix, f, in, fk_pos, fk_rot, fk_vw, blk_list, blk_count, act_list, act_count, act_raw,
                  vb, GRID_FWD_CTAS);
    return;
  }
  vb -= GRID_FWD_CTAS;
  if (vb < shell_ctas_per_env * k.B) {
    // Face cells of unlisted blocks are gathered only by particles whose stencil leaves the grid (clamped / wrapped
    // indices); P2G raises blk_count[1] when it meets one.  An in-range stencil reads only nodes its own P2G scattered
    // into, i.e. cells of listed blocks (block job).
    if (blk_count[1] == 0) return;
    const int env = vb / shell_ctas_per_env;
    grid_shell_body(k, grid_in, grid_out, f, in, fk_pos, fk_rot, fk_vw, act_list, act_count, act_raw, blk_flag, env,
                    vb - env * shell_ctas_per_env);
    return;
  }
  vb -= shell_ctas_per_env * k.B;
  grid_clear_body(k, clr.grid, clr.blk_list, clr.blk_count, clr.shell_ctas_per_env, vb, clr.ctas);
}
static inline int shell_ctas_of(const MpmConst& k) { return cdiv(2 * (k.rx * k.ry + k.rx * k.rz + k.ry * k.rz), 128); }

// clear_grid != null: additionally re-zero `clear_grid` from the block list of substep `clear_substep` (same launch).
// lists_ready: the P2G kernel already appended the touched blocks to this substep's list (warp-local kernels);
// otherwise the block marks are compacted first.
void launch_grid_fwd(const MpmConst& k, float4* grid_in, float4* grid_out, const long long* grid_fix, int substep,
                     const ud_mpm_state* in, const MpmWs& ws, cudaStream_t st, float4* clear_grid, int clear_substep,
                     bool lists_ready) {
  KScope ks_(KC_GRID, st, lists_ready ? 1 : 2);
  int32_t* al = (grid_out != grid_in && ws.act_list) ? ws.act_list + (size_t)(substep - ws.sub0) * k.B * k.G : nullptr;
  int32_t* ac = al ? ws.act_count + substep : nullptr;
  float4* ar = al ? ws.act_raw + (size_t)(substep - ws.sub0) * k.B * k.G : nullptr;
  const int total = k.B * k.nbx * k.nby * k.nbz;
  int32_t* bl = ws.blk_list + (size_t)(substep % ws.blk_nbuf) * total;
  int32_t* bc = ws.blk_count + 2 * substep;
  if (!lists_ready) k_blk_compact<<<cdiv(total, 128), 128, 0, st>>>(total, ws.blk_flag, bl, bc, substep + 1);
  const int sc = shell_ctas_of(k);
  GridClearArgs clr = {nullptr, nullptr, nullptr, 0, sc};
  if (clear_grid) {
    clr.grid = clear_grid;
    clr.blk_list = ws.blk_list + (size_t)(clear_substep % ws.blk_nbuf) * total;
    clr.blk_count = ws.blk_count + 2 * clear_substep;
    clr.ctas = sc * k.B + 148 * 2;
  }
  // in-place (forward) mode the raw value of an empty cell is still there when the shell job reads it
  auto kern = k.sdf_kind == UD_SDF_BOX ? (k.pos_control ? k_grid_fwd<UD_SDF_BOX, 1> : k_grid_fwd<UD_SDF_BOX, 0>)
                                       : (k.pos_control ? k_grid_fwd<UD_SDF_CONTAINER, 1> : k_grid_fwd<UD_SDF_CONTAINER, 0>);
  kern<<<GRID_FWD_CTAS + sc * k.B + clr.ctas, 128, 0, st>>>(k, grid_in, grid_out, const_cast<long long*>(grid_fix), substep, *in,
                                                          ws.fk_pos, ws.fk_rot, ws.fk_vw, bl, bc, al, ac, ar, ws.blk_flag, sc, clr);
}

// ================================================================================================
// Adjoint of the grid update.  One thread per cell with mass and a non-zero incoming cotangent:
// recompute the forward chain in float keeping the velocity that enters each primitive, then reverse:
//   boundary + ground friction : forward-mode Dual<4> over (v, state.friction)  (tiny function)
//   primitives, last to first  : hand-written reverse (collide_cell_bwd / position_control_cell_bwd),
//                                only for primitives whose influence on the cell is >= 1e-12
//   normalise                  : g_p = g_v / m ,  g_m = -g_v . p / m^2
// Per-cell (g_momentum, g_mass) overwrite ggrid; primitive / friction cotangents are warp-reduced and
// accumulated per env.
// ================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}

// Sums each of v[0..N) over the 32 lanes; lane j < N returns the total of v[j] (lanes >= N: 0).  A transposed butterfly:
// every step halves the values a lane carries, 31 shuffles in all instead of 5 per value, and the totals end up one per
// lane, so the N accumulations that follow are ONE atomic instruction of N lanes instead of N by lane 0.
template <int N>
__device__ __forceinline__ float warp_sum_transposed(const float (&vin)[N]) {
  static_assert(N <= 32, "one value per lane");
  const int lane = threadIdx.x & 31;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = i < N ? vin[i] : 0.f;
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float keep = up ? v[i + h] : v[i], send = up ? v[i] : v[i + h];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

constexpr int GRID_BWD_CTAS = 148 * 8;
template <int KIND, int PC>   // compile-time SDF kind / position control: see k_grid_fwd
__global__ void __launch_bounds__(128)
k_grid_bwd(MpmConst k_, const float4* __restrict__ act_raw, float4* __restrict__ ggrid, int f,
           ud_mpm_state in, const float* __restrict__ fk_pos, const float* __restrict__ fk_rot,
           const float* __restrict__ fk_vw, float* __restrict__ g_fk_pos, float* __restrict__ g_fk_rot,
           float* __restrict__ g_fk_v, float* __restrict__ g_scal, const int32_t* __restrict__ act_list,
           const int32_t* __restrict__ act_count, GridClearArgs clr) {
  MpmConst k = k_;
  k.sdf_kind = KIND;
  k.pos_control = PC;
  // CTAs beyond the persistent ones re-zero the OTHER cotangent grid (the one G2P^T of the next, earlier substep
  // scatters into) from the block list of substep f + 1, whose P2G^T has finished: no separate clear launch.
  if (blockIdx.x >= GRID_BWD_CTAS) {
    grid_clear_body(k, clr.grid, clr.blk_list, clr.blk_count, clr.shell_ctas_per_env, blockIdx.x - GRID_BWD_CTAS, clr.ctas);
    return;
  }
  // Persistent grid-stride loop over the cells the recompute pass listed (cells with mass + empty shell cells under
  // a primitive): ~5 % of the grid.  Every other cell keeps the zero the memset gave it.  Whole warps iterate
  // together (the warp reductions below need all lanes).
  const int count = *act_count;
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (GRID_BWD_CTAS * blockDim.x) >> 5;
  for (int wbase = warp0 * 32; wbase < count; wbase += nwarps * 32) {
  const bool live = wbase + lane < count;
  const size_t idx = live ? (size_t)act_list[wbase + lane] : 0;
  const int env = (int)(idx / k.G);
  const int c = (int)(idx - (size_t)env * k.G);
  float4 g = live ? act_raw[wbase + lane] : make_float4(0.f, 0.f, 0.f, 0.f);   // raw {p, m} of the listed cell (coalesced)
  float4 gv4 = ggrid[idx];
  int ck = c % k.rz, cj = (c / k.rz) % k.ry, ci = c / (k.rz * k.ry);
  bool work = live && (gv4.x != 0.f || gv4.y != 0.f || gv4.z != 0.f);
  if (live && !work) ggrid[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!__any_sync(0xffffffffu, work)) continue;  // warp-uniform
  // per-env reductions: the 32 listed cells of a warp almost always belong to one env (warp sum + one atomic);
  // a warp that straddles two envs (or holds dead lanes) lets every lane add its own contribution
  const bool uniform = __match_any_sync(0xffffffffu, live ? env : -1) == 0xffffffffu;
  auto reduce_add = [&](float val, float* addr) {
    if (uniform) {
      const float tot = warp_sum(val);
      if (lane == 0 && tot != 0.f) atomicAdd(addr, tot);
    } else if (val != 0.f) {
      atomicAdd(addr, val);
    }
  };
  const bool has_mass = g.w > 0.f;
  const float gpos[3] = {(float)ci * k.dx, (float)cj * k.dx, (float)ck * k.dx};
  const float sfric = in.friction[env];
  // ---- forward recompute, remembering the velocity entering each primitive
  float v[3] = {0.f, 0.f, 0.f}, vin[UD_MAX_PRIM][3];
  unsigned act = 0;
  if (work) {
    v[0] = (has_mass ? g.x / g.w : g.x) + k.gdt[0];
    v[1] = (has_mass ? g.y / g.w : g.y) + k.gdt[1];
    v[2] = (has_mass ? g.z / g.w : g.z) + k.gdt[2];
#pragma unroll 1   // one copy of the collider code (vin goes to local memory: 3 stores per primitive)
    for (int q = 0; q < UD_MAX_PRIM; ++q) {
      if (q >= k.n_prim) break;
      PrimIn<float> pf;
      load_prim_f<float>(k, in, fk_pos, fk_rot, fk_vw, env, q, f, pf);
      vin[q][0] = v[0];
      vin[q][1] = v[1];
      vin[q][2] = v[2];
      if (!prim_active(k, gpos, pf)) continue;
      act |= 1u << q;
      if (k.pos_control) position_control_cell(k.sdf_kind, k.dt, gpos, pf, v);
      else collide_cell(k.sdf_kind, k.dt, gpos, pf, v);
    }
  }
  // ---- reverse of ground friction + boundary
  float gv[3] = {0.f, 0.f, 0.f}, gsf = 0.f;
  if (work) {
    typedef Dual<4> D4;
    D4 dv[3], sf;
    for (int i = 0; i < 3; ++i) {
      dv[i].v = v[i];
      for (int j = 0; j < 4; ++j) dv[i].d[j] = (i == j) ? 1.f : 0.f;
    }
    sf.v = sfric;
    for (int j = 0; j < 4; ++j) sf.d[j] = (j == 3) ? 1.f : 0.f;
    cell_ground_boundary<D4>(k, ci, cj, ck, sf, dv);
    for (int j = 0; j < 3; ++j) gv[j] = gv4.x * dv[0].d[j] + gv4.y * dv[1].d[j] + gv4.z * dv[2].d[j];
    gsf = gv4.x * dv[0].d[3] + gv4.y * dv[1].d[3] + gv4.z * dv[2].d[3];
  }
  reduce_add(gsf, &g_scal[env * GS_STRIDE + GS_FRICTION]);
  // ---- reverse of the primitives, last to first
  for (int q = k.n_prim - 1; q >= 0; --q) {
    bool mine = work && ((act >> q) & 1u);
    if (!__any_sync(0xffffffffu, mine)) continue;
    PrimGrad pg;
#pragma unroll
    for (int i = 0; i < PRIM_NIN; ++i) pg.g[i] = 0.f;
    if (mine) {
      PrimIn<float> pf;
      load_prim_f<float>(k, in, fk_pos, fk_rot, fk_vw, env, q, f, pf);
      float vq[3] = {0.f, 0.f, 0.f}, gnew[3];
#pragma unroll
      for (int qq = 0; qq < UD_MAX_PRIM; ++qq)
        if (qq == q) {
          vq[0] = vin[qq][0];
          vq[1] = vin[qq][1];
          vq[2] = vin[qq][2];
        }
      if (k.pos_control) position_control_cell_bwd(k.sdf_kind, k.dt, gpos, pf, gv, gnew, pg);
      else collide_cell_bwd(k.sdf_kind, k.dt, gpos, pf, vq, gv, gnew, pg);
      gv[0] = gnew[0];
      gv[1] = gnew[1];
      gv[2] = gnew[2];
    }
    size_t t = (size_t)env * k.n_prim + q;
    float* tp = g_fk_pos + t * (k.S + 1) * 3;
    float* tr = g_fk_rot + t * (k.S + 1) * 4;
    auto dst_of = [&](int j) -> float* {   // accumulator of cotangent j of primitive q (PRIM_NIN order)
      if (j < 3) return &tp[f * 3 + j];
      if (j < 7) return &tr[f * 4 + (j - 3)];
      if (j < 10) return &tp[(f + 1) * 3 + (j - 7)];
      if (j < 14) return &tr[(f + 1) * 4 + (j - 10)];
      if (j < 17) return &g_scal[env * GS_STRIDE + GS_PRIM + q * GS_PRIM_STRIDE + (j - 14)];
      if (j < 18) return &g_scal[env * GS_STRIDE + GS_PRIM + q * GS_PRIM_STRIDE + 3];
      return &g_fk_v[(t * k.S + f) * 3 + (j - 18)];
    };
    if (uniform) {   // one env in the warp: lane j ends up with the warp total of cotangent j and adds it
      const float tot = warp_sum_transposed<PRIM_NIN>(pg.g);
      if (lane < PRIM_NIN && tot != 0.f) atomicAdd(dst_of(lane), tot);
    } else {
#pragma unroll
      for (int j = 0; j < PRIM_NIN; ++j)
        if (pg.g[j] != 0.f) atomicAdd(dst_of(j), pg.g[j]);
    }
  }
  // ---- reverse of v = p / m + dt g
  if (work) {
    float im = has_mass ? 1.f / g.w : 1.f;  // where(m > 0, p / m, p): the empty branch passes the cotangent through
    float gm = has_mass ? -(gv[0] * g.x + gv[1] * g.y + gv[2] * g.z) * im * im : 0.f;
    ggrid[idx] = make_float4(gv[0] * im, gv[1] * im, gv[2] * im, gm);
  }
  }  // listed cells
}

// ggrid: the cotangent grid G2P^T(substep) scattered into, reversed in place.  clear_grid != null: additionally
// re-zero `clear_grid` from the block list of `clear_substep` in the same launch.
void launch_grid_bwd(const MpmConst& k, int substep, const ud_mpm_state* in, const MpmWs& ws, cudaStream_t st, float4* ggrid,
                     float4* clear_grid, int clear_substep) {
  KScope ks_(KC_GRID_BWD, st);
  const int total = k.B * k.nbx * k.nby * k.nbz;
  const int sc = shell_ctas_of(k);
  GridClearArgs clr = {nullptr, nullptr, nullptr, 0, sc};
  if (clear_grid) {
    clr.grid = clear_grid;
    clr.blk_list = ws.blk_list + (size_t)(clear_substep % ws.blk_nbuf) * total;
    clr.blk_count = ws.blk_count + 2 * clear_substep;
    clr.ctas = sc * k.B + 148 * 2;
  }
  auto kern = k.sdf_kind == UD_SDF_BOX ? (k.pos_control ? k_grid_bwd<UD_SDF_BOX, 1> : k_grid_bwd<UD_SDF_BOX, 0>)
                                       : (k.pos_control ? k_grid_bwd<UD_SDF_CONTAINER, 1> : k_grid_bwd<UD_SDF_CONTAINER, 0>);
  kern<<<GRID_BWD_CTAS + clr.ctas, 128, 0, st>>>(k, ws.act_raw + (size_t)(substep - ws.sub0) * k.B * k.G, ggrid, substep, *in, ws.fk_pos, ws.fk_rot, ws.fk_vw, ws.g_fk_pos,
                                                 ws.g_fk_rot, ws.g_fk_v, ws.g_scal,
                                                 ws.act_list + (size_t)(substep - ws.sub0) * k.B * k.G, ws.act_count + substep, clr);
}

// ------------------------------------------------------------------------------------------------
// Reverse of the primitive prologue: copy_frame^T, FK chain^T (position clip masks, quaternion
// products), set_action^T, action clip mask.  One thread per (env, primitive).
// Writes g_prim_in = {g position[0] (3), g rotation[0] (4), g action_scale (6)} and g_act (6).
// ------------------------------------------------------------------------------------------------
// One WARP per (env, primitive): the S+1 rows of the position / rotation cotangent tables are staged into shared
// memory with parallel loads (the per-row output-table cotangents are added on the way, the (v, w) row cotangents are
// warp-reduced), then lane 0 walks the short sequential chain on shared memory.  The previous one-thread version did
// ~400 dependent global read-modify-writes (117 us for S = 16).
constexpr int FKB_BLOCK = 128;
__global__ void __launch_bounds__(FKB_BLOCK)
k_fk_bwd(MpmConst k, ud_mpm_state in, const float* __restrict__ action, ud_mpm_state gout,
         const float* __restrict__ fk_pos, const float* __restrict__ fk_rot,
         const float* __restrict__ fk_vw, const float* __restrict__ fk_act,
         float* __restrict__ g_fk_pos, float* __restrict__ g_fk_rot,
         const float* __restrict__ g_fk_v, float* __restrict__ g_prim_in,
         float* __restrict__ g_act) {
  extern __shared__ float fksm[];
  const int t = blockIdx.x, lane = threadIdx.x;
  const int env = t / k.n_prim, q = t % k.n_prim;
  const int S = k.S;
  float* gtp = fksm;                  // [(S+1)*3]
  float* gtr = fksm + (S + 1) * 3;    // [(S+1)*4]
  float* jac = fksm + (S + 1) * 7;    // [S-1][4][8]: d qmul(dq, rot[f]) / d(dq, rot[f]) of every row
  const ud_primitive& go = gout.prim[q];
  const float* tp = fk_pos + (size_t)t * (S + 1) * 3;
  const float* tr = fk_rot + (size_t)t * (S + 1) * 4;
  const float* gp_in = g_fk_pos + (size_t)t * (S + 1) * 3;
  const float* gr_in = g_fk_rot + (size_t)t * (S + 1) * 4;
  // The Jacobians of the rotation chain's rows (forward mode over the 8 inputs of qmul: ~1 500 dependent instructions
  // each, IEEE divisions) do not depend on each other: one row per thread, all of them at once.  The sequential part
  // that is left for thread 0 is a 4x8 matrix-vector product per row.  (Round 2 until its last session evaluated them
  // inside the sequential loop: 80 us per call at S = 16 and ~4 us for every further row.)
  {
    float vwr[3], dq0[4];
#pragma unroll
    for (int j = 0; j < 3; ++j) vwr[j] = fk_vw[(size_t)t * 6 + 3 + j];
    w2quat(vwr, dq0);
    for (int f = threadIdx.x; f < S - 1; f += FKB_BLOCK) {
      Dual<8> dqd[4], rd[4], out[4];
      for (int j = 0; j < 4; ++j) {
        dqd[j].v = dq0[j];
        rd[j].v = tr[f * 4 + j];
        for (int i = 0; i < 8; ++i) {
          dqd[j].d[i] = (i == j) ? 1.f : 0.f;
          rd[j].d[i] = (i == 4 + j) ? 1.f : 0.f;
        }
      }
      qmul(dqd, rd, out);
      for (int o = 0; o < 4; ++o)
        for (int i = 0; i < 8; ++i) jac[(f * 4 + o) * 8 + i] = out[o].d[i];
    }
  }
  if (threadIdx.x >= 32) {   // warps 1.. only computed Jacobians
    __syncthreads();
    return;
  }
  // rows 1..S-1 of the output tables are internal rows 1..S-1; output row 0 = internal row S-1 (copy_frame)
  for (int e = lane; e < (S + 1) * 3; e += 32) {
    const int r = e / 3;
    float vsum = gp_in[e];
    if (r >= 1 && r < S && go.position) vsum += go.position[(size_t)env * S * 3 + e];
    gtp[e] = vsum;
  }
  for (int e = lane; e < (S + 1) * 4; e += 32) {
    const int r = e / 4;
    float vsum = gr_in[e];
    if (r >= 1 && r < S && go.rotation) vsum += go.rotation[(size_t)env * S * 4 + e];
    gtr[e] = vsum;
  }
  float gvw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int f = lane; f < S; f += 32) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (go.v) gvw[j] += go.v[((size_t)env * S + f) * 3 + j];
      if (go.w) gvw[3 + j] += go.w[((size_t)env * S + f) * 3 + j];
      gvw[j] += g_fk_v[((size_t)t * S + f) * 3 + j];
    }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) gvw[j] += __shfl_xor_sync(0xffffffffu, gvw[j], off);
  __syncthreads();   // the staged tables and every row's Jacobian are in shared memory
  if (lane != 0) return;
  // row S is the clamped alias of row S-1; output row 0 lands on internal row S-1
  for (int j = 0; j < 3; ++j) gtp[(S - 1) * 3 + j] += gtp[S * 3 + j] + (go.position ? go.position[(size_t)env * S * 3 + j] : 0.f);
  for (int j = 0; j < 4; ++j) gtr[(S - 1) * 4 + j] += gtr[S * 4 + j] + (go.rotation ? go.rotation[(size_t)env * S * 4 + j] : 0.f);
  float vw[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) vw[j] = fk_vw[(size_t)t * 6 + j];
  float dq[4];
  w2quat(vw + 3, dq);
  float gdq[4] = {0.f, 0.f, 0.f, 0.f};
  for (int f = S - 2; f >= 0; --f) {
    // pos[f+1] = clip(pos[f] + v, -2, 2)
    for (int j = 0; j < 3; ++j) {
      float u = tp[f * 3 + j] + vw[j];
      float gnext = clip_grad(u, -2.f, 2.f) * gtp[(f + 1) * 3 + j];
      gtp[f * 3 + j] += gnext;
      gvw[j] += gnext;
    }
    // rot[f+1] = qmul(dq, rot[f]): transposed Jacobian (computed above) times the cotangent of row f+1
    for (int i = 0; i < 4; ++i) {
      float a = 0.f, b = 0.f;
      for (int o = 0; o < 4; ++o) {
        a += gtr[(f + 1) * 4 + o] * jac[(f * 4 + o) * 8 + i];
        b += gtr[(f + 1) * 4 + o] * jac[(f * 4 + o) * 8 + 4 + i];
      }
      gdq[i] += a;
      gtr[f * 4 + i] += b;
    }
  }
  {  // dq = w2quat(w row)
    Dual<3> wd[3], out[4];
    for (int j = 0; j < 3; ++j) {
      wd[j].v = vw[3 + j];
      for (int i = 0; i < 3; ++i) wd[j].d[i] = (i == j) ? 1.f : 0.f;
    }
    w2quat(wd, out);
    for (int i = 0; i < 3; ++i)
      for (int o = 0; o < 4; ++o) gvw[3 + i] += gdq[o] * out[o].d[i];
  }
  float* pi = g_prim_in + (size_t)t * 16;
  for (int j = 0; j < 3; ++j) pi[j] = gtp[j];
  for (int j = 0; j < 4; ++j) pi[3 + j] = gtr[j];
  const ud_primitive& pin = in.prim[q];
  for (int j = 0; j < 6; ++j) {
    float a = fk_act[(size_t)t * 6 + j];
    float scale = pin.action_scale[env * 6 + j];
    float ga = gvw[j] * scale / (float)S + (go.action_buffer ? go.action_buffer[env * 6 + j] : 0.f);
    pi[7 + j] = gvw[j] * a / (float)S + (go.action_scale ? go.action_scale[env * 6 + j] : 0.f);
    float raw = action[(size_t)env * 6 * k.n_prim + 6 * q + j];
    g_act[(size_t)t * 6 + j] = clip_grad(raw, -1.f, 1.f) * ga;
  }
  for (int j = 13; j < 16; ++j) pi[j] = 0.f;
}

void launch_fk_bwd(const MpmConst& k, const ud_mpm_state* in, const float* action,
                   const ud_mpm_state* gout, const MpmWs& ws, cudaStream_t st) {
  int n = k.B * k.n_prim;
  if (n == 0) return;
  KScope ks_(KC_FK, st);
  const size_t smem = sizeof(float) * (7 * (size_t)(k.S + 1) + 32 * (size_t)(k.S > 1 ? k.S - 1 : 0));   // 20.6 KB at S = 133
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_fk_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
  k_fk_bwd<<<n, FKB_BLOCK, smem, st>>>(k, *in, action, *gout, ws.fk_pos, ws.fk_rot, ws.fk_vw, ws.fk_act, ws.g_fk_pos,
                                       ws.g_fk_rot, ws.g_fk_v, ws.g_prim_in, ws.g_act);
}

}  // namespace ud
