// Internal (non-ABI) declarations shared by the MPM translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/unidom_b200.h"
#include "mpm_particle.cuh"

namespace ud {

// Sorted particle state is an array of warp tiles (AoSoA): tile t = gp >> 5 holds 32 consecutive particles of the
// per-frame sort, gp = env * n_pad + slot (every env starts on a tile boundary).  Inside a tile the components are
// grouped in QUADS: float4 number (t * NQ + q) * 32 + lane holds components 4q..4q+3 of particle `lane`, so a warp
// moves 4 components of its 32 particles with ONE 512-byte LDG.128 / STG.128 and every address is
// tile base + immediate.  State tile (NQ = 6):
//   0..2 x, 3..5 v, 6..14 C (row-major), 15..23 F (row-major)
constexpr int PS_X = 0, PS_V = 3, PS_C = 6, PS_F = 15, PS_NCOMP = 24, PS_NQ = 6;
// SVD of F1 kept by the recompute pass for the adjoint: U (row-major 9), s (3), Vt (row-major 9), 3 pad (NQ = 6)
constexpr int SV_U = 0, SV_S = 9, SV_VT = 12, SV_NCOMP = 24, SV_NQ = 6;
// forward-only warm-start buffer: Vt (9) + 3 pad (NQ = 3)
constexpr int VT_NCOMP = 12, VT_NQ = 3;

UD_DEV size_t quad_index(int gp, int q, int nq) { return ((size_t)(gp >> 5) * nq + q) * 32 + (gp & 31); }
UD_DEV size_t comp_index(int gp, int c, int nq) { return quad_index(gp, c >> 2, nq) * 4 + (c & 3); }

// components [C0, C0+NC) of particle gp from a tile array with NQ quads: one 16-byte load per covering quad
template <int C0, int NC, int NQ, bool CG = false>
__device__ __forceinline__ void load_comps(const float* __restrict__ base, int gp, float* out) {
  const float4* t = reinterpret_cast<const float4*>(base) + quad_index(gp, 0, NQ);
  constexpr int Q0 = C0 / 4, Q1 = (C0 + NC - 1) / 4;
#pragma unroll
  for (int q = Q0; q <= Q1; ++q) {
    float4 v;
    if (CG) {   // "reload": a volatile L1-cached load the compiler can neither merge with an earlier one nor hoist
      asm volatile("ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(t + q * 32));
    } else {
      v = t[q * 32];
    }
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = q * 4 + i;
      if (c >= C0 && c < C0 + NC) out[c - C0] = e[i];
    }
  }
}
// stores components [C0, C0+NC): whole quads as one 16-byte store, partially covered quads component by component
template <int C0, int NC, int NQ>
__device__ __forceinline__ void store_comps(float* __restrict__ base, int gp, const float* in) {
  float4* t = reinterpret_cast<float4*>(base) + quad_index(gp, 0, NQ);
  constexpr int Q0 = C0 / 4, Q1 = (C0 + NC - 1) / 4;
#pragma unroll
  for (int q = Q0; q <= Q1; ++q) {
    const bool full = q * 4 >= C0 && q * 4 + 3 < C0 + NC;
    if (full) {
      t[q * 32] = make_float4(in[q * 4 - C0], in[q * 4 + 1 - C0], in[q * 4 + 2 - C0], in[q * 4 + 3 - C0]);
    } else {
      float* f = reinterpret_cast<float*>(t + q * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = q * 4 + i;
        if (c >= C0 && c < C0 + NC) f[i] = in[c - C0];
      }
    }
  }
}
// Block flags (ws.blk_flag, zeroed once per call): flag == substep + 1 <=> the block is on that substep's list.  The
// warp-local P2G kernels stamp and list themselves; the round-1 CTA kernels store BLK_MARK_CTA and k_blk_compact turns
// it into the stamp while listing.
constexpr int BLK_MARK_CTA = 0x40000000;
// the SVD warm-start chain restarts from V = I every SVD_RESTART substeps (bounds rounding drift of V)
constexpr int SVD_RESTART = 16;
// deterministic P2G: values are accumulated as round(v * 2^52) in int64 (range +-2048, resolution 2.2e-16);
// integer addition is associative, so the result is independent of the order the REDs arrive in
constexpr double FIX_SCALE = 4503599627370496.0, FIX_INV = 1.0 / 4503599627370496.0;
// per-env scalar cotangent accumulators (bwd)
//   0 friction, 1 mu, 2 lamda, then per primitive q: 3+4q .. : size(3), friction(1)
constexpr int GS_FRICTION = 0, GS_MU = 1, GS_LAMDA = 2, GS_PRIM = 3, GS_PRIM_STRIDE = 4;
constexpr int GS_STRIDE = GS_PRIM + GS_PRIM_STRIDE * UD_MAX_PRIM;

struct MpmWs {
  // binning
  int32_t* inv_perm;    // [N] sorted slot of original particle p (inverse of perm)
  int32_t* chunk_sum;   // [B * ceil(NK / 1024)] totals of the 1024-key chunks of the count scan
  int32_t* keys;        // [N]
  int32_t* tmp_idx;     // [N]
  int32_t* perm;        // [N] sorted slot -> original particle index (per env)
  int32_t* cell_start;  // [B*(NK+1)]
  int32_t* cursor;      // [B*NK + 1]
  int32_t* big_list;    // [2 * (N / RANK_BIG + 1)]
  int32_t* mat_s;       // [N] material in sorted order
  float* h_s;           // [N] hardness in sorted order
  // state
  float* ps;            // fwd: [24*N_pad]; bwd: [(S+1)*24*N_pad] start-of-substep states (tiles, see above)
  float4* grid_raw;     // [2][B*G] scattered (p, m), used alternately: the grid launch of substep f re-zeroes the other one
  float4* grid_out;     // fwd: == grid_raw; bwd: [S*B*G] updated velocities
  long long* grid_fix;  // deterministic P2G only: [B*G*4] 64-bit fixed-point accumulators of one substep
  int32_t* blk_flag;    // [B*nbx*nby*nbz] 4x4x4 grid blocks that P2G scattered into this substep
  int32_t* blk_list;    // [blk_nbuf][B*nbx*nby*nbz] the marked blocks, compacted; slot = substep % blk_nbuf
  int blk_nbuf;         // fwd: 2 (k_grid_clear reads the previous substep's list); bwd: S (the reverse pass re-zeroes
                        // the cotangent grid block by block)
  int32_t* blk_count;   // [S][2] per substep: number of listed blocks; 1 when some particle's stencil leaves the grid
                        // (only then does anything gather a face cell of a block nobody scattered into: the shell
                        // job of the grid update and the face-cell re-zeroing run only for such substeps)
  float* vt_roll;       // fwd only: [12*N_pad] V^T of the previous substep's SVD (warm start)
  int32_t* act_list;    // bwd only: [S][B*G] cells listed by the recompute pass for the grid adjoint
  float4* act_raw;      // bwd only: [S][B*G] their raw (p, m), in list order (the raw grids do not outlive their substep)
  // windowed adjoint only (ud_mpm_step_bwd_windowed): K-spaced checkpoints and the checkpoint pass's scratch
  float* ckpt_ps;       // [n_win][24*N_pad] start state of substep w*window (slot 0 = the gathered step input)
  float* ckpt_vt;       // [n_win][12*N_pad] V^T entering that substep (SVD warm start)
  float* run_ps;        // [24*N_pad] in-place state of the checkpoint pass
  float* run_vt;        // [12*N_pad]
  float4* run_grid;     // [2][B*G]
  int sub0;             // first substep of the window the per-substep arrays (act_list) start at; 0 outside the windowed adjoint
  int32_t* act_count;   // bwd only: [S]
  float* svd_s;         // bwd only: [S*24*N_pad] SVD of F1 per substep (written by the recompute P2G)
  float* fk_pos;        // [B*P*(S+1)*3] (row S = clamp copy of row S-1)
  float* fk_rot;        // [B*P*(S+1)*4]
  float* fk_vw;         // [B*P*6] per-substep (v,w) row
  float* fk_act;        // [B*P*6] clipped action
  float* jrows;         // [B*S*9] rows i<3 of C' of original particles 0..2 (J update, :327)
  // adjoint
  float* gs;            // [24*N_pad] cotangent tiles (sorted order)
  float4* ggrid;        // [2][B*G] cotangent grids, used alternately: substep f scatters into / gathers from slot f & 1
                        // while the grid adjoint's launch re-zeroes the other one block by block
  float* g_fk_pos;      // [B*P*(S+1)*3]
  float* g_fk_rot;      // [B*P*(S+1)*4]
  float* g_fk_v;        // [B*P*S*3] (position-control rows)
  float* g_scal;        // [B*GS_STRIDE]
  float* g_prim_in;     // [B*P*16]: gpos0(3) grot0(4) gscale(6) pad(3)
  float* g_act;         // [B*P*6] cotangent of the (unclipped) action
  float* norm2;         // [B*2] (state norm^2, action norm^2)
  size_t bytes;
};

// Kernel classes for the optional per-kernel CUDA-event timing (ud_timing_*).
enum KClass { KC_SORT = 0, KC_GATHER, KC_FK, KC_P2G, KC_GRID, KC_G2P, KC_UNSORT, KC_G2P_BWD, KC_GRID_BWD,
              KC_P2G_BWD, KC_FINISH_BWD, KC_MEMSET, KC_CLOTH_FWD, KC_CLOTH_BWD, KC_REWARD, KC_APG, KC_COUNT };
// RAII scope: counts launches and (when timing is enabled) brackets them with events on `st`.
struct KScope {
  int cls;
  cudaStream_t st;
  int slot;
  KScope(int cls, cudaStream_t st, int n_launches = 1);
  ~KScope();
};

// A/B switches (ud_tuning_set): 1 = the fast path (default)
int tuning_sort();   // 0: identity permutation instead of the per-frame binning
int tuning_stage();  // 0: 27 vector REDs per particle straight to HBM instead of the shared-memory staged scatter

// Host helpers (abi.cu)
int set_error(int code, const char* what);   // records the message ud_last_error() returns (thread-local); returns code
bool mpm_fold_constants(const ud_mpm_params* p, MpmConst* k);
size_t mpm_carve(const ud_mpm_params* p, const MpmConst& k, bool bwd, void* base, MpmWs* ws, int window = 0);

// Launchers implemented in mpm_particles.cu
void launch_sort(const MpmConst& k, const float* x_aos, const MpmWs& ws, int32_t* out_base, cudaStream_t st);
void launch_gather_state(const MpmConst& k, const ud_mpm_state* in, const int32_t* material, const float* h,
                         const MpmWs& ws, float* ps_slot, cudaStream_t st);
// vt_in: warm start of the SVD (null: cold).  It is the forward's V^T buffer when svd_out is null or vt_in_is_vt is
// set, else the previous substep's SVD tile.
void launch_p2g(const MpmConst& k, const float* ps_in, float* ps_out, float4* grid, const float* mu_s,
                const float* la_s, const float* vt_in, float* vt_out, float* svd_out, int substep, const MpmWs& ws,
                cudaStream_t st, bool vt_in_is_vt = false);
bool p2g_lists_blocks();   // the P2G kernels in use list the touched grid blocks themselves (no compaction pass)
void launch_g2p(const MpmConst& k, const float* ps_in, float* ps_out, const float4* grid, int substep,
                const MpmWs& ws, cudaStream_t st);
void launch_unsort_state(const MpmConst& k, const float* ps_slot, const float* J_in, const MpmWs& ws,
                         ud_mpm_state* out, cudaStream_t st);
void launch_gather_cot(const MpmConst& k, const ud_mpm_state* gout, const MpmWs& ws, cudaStream_t st);
void launch_g2p_bwd(const MpmConst& k, const float* ps_in, const float4* grid_out, const MpmWs& ws,
                    cudaStream_t st, float4* ggrid);
void launch_p2g_bwd(const MpmConst& k, const float* ps_in, const float* svd_in, const float* mu_s,
                    const float* la_s, bool first_substep, const MpmWs& ws, cudaStream_t st, const float4* ggrid);
void launch_finish_bwd(const MpmConst& k, const ud_mpm_state* in, const ud_mpm_state* gout,
                       ud_mpm_state* gin, const float* action, float* gaction, const MpmWs& ws,
                       cudaStream_t st);

// Launchers implemented in mpm_grid.cu (compiled with --fmad=false)
void launch_fk_fwd(const MpmConst& k, const ud_mpm_state* in, const float* action, ud_mpm_state* out,
                   const MpmWs& ws, cudaStream_t st);
// grid_fix != null (deterministic P2G): the raw {p,m} is first converted from the fixed-point accumulators
// into grid_in (which the adjoint reads later), then updated into grid_out.
void launch_grid_fwd(const MpmConst& k, float4* grid_in, float4* grid_out, const long long* grid_fix, int substep,
                     const ud_mpm_state* in, const MpmWs& ws, cudaStream_t st, float4* clear_grid = nullptr,
                     int clear_substep = 0, bool lists_ready = false);
void launch_grid_bwd(const MpmConst& k, int substep, const ud_mpm_state* in, const MpmWs& ws, cudaStream_t st, float4* ggrid,
                     float4* clear_grid, int clear_substep);
void launch_fk_bwd(const MpmConst& k, const ud_mpm_state* in, const float* action,
                   const ud_mpm_state* gout, const MpmWs& ws, cudaStream_t st);

}  // namespace ud
