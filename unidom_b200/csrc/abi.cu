// C ABI of unidom_b200 (see include/unidom_b200.h): validation, constant folding, workspace carving
// and the per-step launch sequences.  Nothing here synchronises or allocates.
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "mpm_internal.h"

namespace ud {

// ---- instrumentation ---------------------------------------------------------------------------
static std::atomic<uint64_t> g_launches{0};
static bool g_timing = false;
struct TimingRec { int cls; int n; cudaEvent_t a, b; };
static std::vector<TimingRec> g_recs;      // in flight since the last collect
static std::vector<cudaEvent_t> g_pool;    // recycled events
static std::mutex g_tmu;

static cudaEvent_t ev_get() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
KScope::KScope(int cls_, cudaStream_t st_, int n) : cls(cls_), st(st_), slot(-1) {
  g_launches += (uint64_t)n;
  if (!g_timing) return;
  std::lock_guard<std::mutex> lk(g_tmu);
  TimingRec r; r.cls = cls; r.n = n; r.a = ev_get(); r.b = ev_get();
  cudaEventRecord(r.a, st);
  slot = (int)g_recs.size();
  g_recs.push_back(r);
}
KScope::~KScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_tmu);
  cudaEventRecord(g_recs[slot].b, st);
}

static int g_svd_warm = 1, g_sort = 1, g_stage = 1, g_mark = 2;
static thread_local const char* g_last_error = "";
int set_error(int code, const char* what) {
  g_last_error = what;
  return code;
}
static int fail(int code, const char* what) { return set_error(code, what); }

int cloth_tuning_cta_nodes(int v);   // cloth.cu
int tuning_sort() { return g_sort; }
int tuning_stage() { return g_stage; }
int tuning_warp(int v);   // mpm_particles.cu
int tuning_pers(int v);
int tuning_apg_rs(int v);

bool mpm_fold_constants(const ud_mpm_params* p, MpmConst* k) {
  if (!p) return false;
  if (p->num_envs < 1 || p->n_particles < 1 || p->steps < 1) return false;
  if (p->res[0] < 1 || p->res[1] < 1 || p->res[2] < 1 || p->n_grid < 1) return false;
  if (p->n_primitive < 0 || p->n_primitive > UD_MAX_PRIM) return false;
  if (p->sdf_kind != UD_SDF_BOX && p->sdf_kind != UD_SDF_CONTAINER) return false;
  const int p2g_mode = p->p2g_mode & ~UD_P2G_LIQUID_FAST;
  if (p2g_mode != UD_P2G_ATOMIC && p2g_mode != UD_P2G_DETERMINISTIC) return false;
  if (!(p->dt > 0) || !(p->dx > 0) || !(p->inv_dx > 0) || !(p->p_mass > 0) || !(p->p_vol > 0)) return false;
  long long N = (long long)p->num_envs * p->n_particles;
  long long G = (long long)p->res[0] * p->res[1] * p->res[2];
  if ((N + 32LL * p->num_envs) > 0x7fffffffLL / 32 || G > 0x7fffffffLL / 4) return false;
  memset(k, 0, sizeof(*k));
  k->B = p->num_envs;
  k->n = p->n_particles;
  k->S = p->steps;
  k->N = (int)N;
  k->n_pad = (p->n_particles + 31) / 32 * 32;
  k->N_pad = k->B * k->n_pad;
  k->G = (int)G;
  k->rx = p->res[0];
  k->ry = p->res[1];
  k->rz = p->res[2];
  k->n_grid = p->n_grid;
  k->nbx = (k->rx + 3) / 4;
  k->nby = (k->ry + 3) / 4;
  k->nbz = (k->rz + 3) / 4;
  k->NK = k->nbx * k->nby * k->nbz * 64;
  k->dt = (float)p->dt;
  k->dx = (float)p->dx;
  k->inv_dx = (float)p->inv_dx;
  k->p_mass = (float)p->p_mass;
  k->c_stress_mul = (float)(-p->dt * p->p_vol * 4);  // mpm_simulator.py:267, folded in double like Python
  k->c_stress_div = (float)(p->dx * p->dx);
  for (int i = 0; i < 3; ++i) k->gdt[i] = (float)p->dt * (float)p->gravity[i];  // :285, f32*f32
  k->sig_lo = (float)(1 - 2.5e-2 * 10);
  k->sig_hi = (float)(1 + 4.5e-3 * 100);
  k->n_prim = p->n_primitive;
  k->sdf_kind = p->sdf_kind;
  k->pos_control = p->use_position_control ? 1 : 0;
  k->p2g_mode = p2g_mode;
  k->liquid_fast = (p->p2g_mode & UD_P2G_LIQUID_FAST) ? 1 : 0;
  k->mark = g_mark;
  return true;
}

static inline void zero_async(void* p, size_t bytes, cudaStream_t st) {
  KScope ks(KC_MEMSET, st);
  cudaMemsetAsync(p, 0, bytes, st);
}

static inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

// Carves `base` (may be null: size query only).  Returns total bytes.
// window > 0: layout of the windowed adjoint (per-substep arrays hold `window` substeps, plus the checkpoints).
size_t mpm_carve(const ud_mpm_params*, const MpmConst& k, bool bwd, void* base, MpmWs* ws, int window) {
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t bytes) {
    void* r = b ? (void*)(b + off) : nullptr;
    off += al(bytes);
    return r;
  };
  const size_t N = k.N, NP = k.N_pad, BG = (size_t)k.B * k.G, P = k.n_prim > 0 ? k.n_prim : 1, S = k.S;
  const size_t W = window > 0 ? (size_t)window : S;   // substeps resident at a time in the per-substep arrays
  MpmWs w;
  memset(&w, 0, sizeof(w));
  w.keys = (int32_t*)take(4 * N);
  w.tmp_idx = (int32_t*)take(4 * N);
  w.perm = (int32_t*)take(4 * N);
  w.inv_perm = (int32_t*)take(4 * N);
  w.cell_start = (int32_t*)take(4 * (size_t)k.B * (k.NK + 1));
  w.cursor = (int32_t*)take(4 * ((size_t)k.B * k.NK + 1));   // [B*NK] + the crowded-cell counter of k_rank
  w.big_list = (int32_t*)take(4 * 2 * (N / 160 + 1));        // (env, key) of the cells k_rank leaves to k_rank_big
  w.chunk_sum = (int32_t*)take(4 * (size_t)k.B * ((k.NK + 1023) / 1024));
  w.mat_s = (int32_t*)take(4 * N);
  w.h_s = (float*)take(4 * N);
  w.fk_pos = (float*)take(4 * (size_t)k.B * P * (S + 1) * 3);
  w.fk_rot = (float*)take(4 * (size_t)k.B * P * (S + 1) * 4);
  w.fk_vw = (float*)take(4 * (size_t)k.B * P * 6);
  w.fk_act = (float*)take(4 * (size_t)k.B * P * 6);
  w.jrows = (float*)take(4 * (size_t)k.B * S * 9);
  w.blk_flag = (int32_t*)take(4 * (size_t)k.B * k.nbx * k.nby * k.nbz);
  w.blk_nbuf = bwd ? k.S : 2;
  w.blk_list = (int32_t*)take(4 * (size_t)w.blk_nbuf * k.B * k.nbx * k.nby * k.nbz);
  w.blk_count = (int32_t*)take(4 * 2 * S);   // [S][2]: {listed blocks, shell job needed}
  if (!bwd) {
    w.ps = (float*)take(4 * (size_t)PS_NCOMP * NP);
    w.vt_roll = (float*)take(4 * (size_t)VT_NCOMP * NP);
    w.grid_raw = (float4*)take(16 * BG * 2);
    w.grid_out = w.grid_raw;
  } else {
    w.ps = (float*)take(4 * (size_t)PS_NCOMP * NP * (W + 1));
    w.grid_raw = (float4*)take(16 * BG * 2);
    w.act_raw = (float4*)take(16 * BG * W);
    w.grid_out = (float4*)take(16 * BG * W);
    w.svd_s = (float*)take(4 * (size_t)SV_NCOMP * NP * W);
    w.act_list = (int32_t*)take(4 * BG * W);
    if (window > 0) {
      const size_t n_win = (S + W - 1) / W;
      w.ckpt_ps = (float*)take(4 * (size_t)PS_NCOMP * NP * n_win);
      w.ckpt_vt = (float*)take(4 * (size_t)VT_NCOMP * NP * n_win);
      w.run_ps = (float*)take(4 * (size_t)PS_NCOMP * NP);
      w.run_vt = (float*)take(4 * (size_t)VT_NCOMP * NP);
      w.run_grid = (float4*)take(16 * BG * 2);
    }
    w.act_count = (int32_t*)take(4 * S);
    w.gs = (float*)take(4 * (size_t)PS_NCOMP * NP);
    w.ggrid = (float4*)take(16 * BG * 2);
    w.g_fk_pos = (float*)take(4 * (size_t)k.B * P * (S + 1) * 3);
    w.g_fk_rot = (float*)take(4 * (size_t)k.B * P * (S + 1) * 4);
    w.g_fk_v = (float*)take(4 * (size_t)k.B * P * S * 3);
    w.g_scal = (float*)take(4 * (size_t)k.B * GS_STRIDE);
    w.g_prim_in = (float*)take(4 * (size_t)k.B * P * 16);
    w.g_act = (float*)take(4 * (size_t)k.B * P * 6);
    w.norm2 = (float*)take(4 * (size_t)k.B * 2);
  }
  if (k.p2g_mode == UD_P2G_DETERMINISTIC) w.grid_fix = (long long*)take(32 * BG);
  w.bytes = off;
  if (ws) *ws = w;
  return off;
}

static bool state_ok(const MpmConst& k, const ud_mpm_state* s) {
  if (!s || !s->x || !s->v || !s->C || !s->F || !s->J || !s->friction || !s->mu || !s->lamda) return false;
  for (int q = 0; q < k.n_prim; ++q) {
    const ud_primitive& p = s->prim[q];
    if (!p.size || !p.friction || !p.softness || !p.position || !p.rotation || !p.v || !p.w ||
        !p.action_buffer || !p.action_scale)
      return false;
  }
  return true;
}

}  // namespace ud

using namespace ud;

extern "C" {

const char* ud_version(void) { return "unidom_b200 0.1 (sm_100a)"; }

uint64_t ud_launch_count(int reset) {
  uint64_t v = g_launches.load();
  if (reset) g_launches = 0;
  return v;
}
void ud_timing_enable(int on) { g_timing = on != 0; }
int ud_tuning_set(const char* name, int value) {
  if (name && !strcmp(name, "svd_warm")) { int o = g_svd_warm; g_svd_warm = value; return o; }
  if (name && !strcmp(name, "sort")) { int o = g_sort; g_sort = value; return o; }
  if (name && !strcmp(name, "stage")) { int o = g_stage; g_stage = value; return o; }
  if (name && !strcmp(name, "mark")) { int o = g_mark; g_mark = value; return o; }
  if (name && !strcmp(name, "warp")) return tuning_warp(value);
  if (name && !strcmp(name, "pers")) return tuning_pers(value);
  if (name && !strcmp(name, "cloth_cta_nodes")) return cloth_tuning_cta_nodes(value);
  if (name && !strcmp(name, "apg_rs")) return tuning_apg_rs(value) + 1;   // (old value + 1: the old value may be -1 = auto)
  return -1;
}
int ud_timing_num_classes(void) { return KC_COUNT; }
const char* ud_timing_class_name(int cls) {
  static const char* names[KC_COUNT] = {"sort", "gather", "fk", "p2g", "grid", "g2p", "unsort", "g2p_bwd",
                                        "grid_bwd", "p2g_bwd", "finish_bwd", "memset", "cloth_fwd", "cloth_bwd", "reward",
                                        "apg"};
  return (cls >= 0 && cls < KC_COUNT) ? names[cls] : "";
}
int ud_timing_collect(double* ms_by_class, int64_t* launches_by_class, int n_classes) {
  if (!ms_by_class || !launches_by_class || n_classes < KC_COUNT) return UD_E_INVALID;
  std::lock_guard<std::mutex> lk(g_tmu);
  for (int i = 0; i < n_classes; ++i) { ms_by_class[i] = 0; launches_by_class[i] = 0; }
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    ms_by_class[r.cls] += ms;
    launches_by_class[r.cls] += r.n;
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  return UD_OK;
}
const char* ud_last_error(void) { return g_last_error; }

size_t ud_mpm_fwd_workspace_bytes(const ud_mpm_params* p) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k)) return 0;
  return mpm_carve(p, k, false, nullptr, nullptr);
}
size_t ud_mpm_bwd_workspace_bytes(const ud_mpm_params* p) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k)) return 0;
  return mpm_carve(p, k, true, nullptr, nullptr);
}
int32_t ud_mpm_num_keys(const ud_mpm_params* p) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k)) return 0;
  return k.NK;
}

int ud_mpm_sort_bins(const ud_mpm_params* p, const float* x, int32_t* out_base, int32_t* out_key,
                     int32_t* out_perm, void* workspace, size_t workspace_bytes, void* stream) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k)) return fail(UD_E_INVALID, "ud_mpm_sort_bins: invalid params");
  if (!x || !out_base || !out_key || !out_perm) return fail(UD_E_INVALID, "ud_mpm_sort_bins: null pointer");
  MpmWs ws;
  size_t need = mpm_carve(p, k, false, workspace, &ws);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255))
    return fail(UD_E_WORKSPACE, "ud_mpm_sort_bins: workspace too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  launch_sort(k, x, ws, out_base, st);
  cudaMemcpyAsync(out_key, ws.keys, 4 * (size_t)k.N, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(out_perm, ws.perm, 4 * (size_t)k.N, cudaMemcpyDeviceToDevice, st);
  if (cudaGetLastError() != cudaSuccess) return fail(UD_E_CUDA, "ud_mpm_sort_bins: launch failed");
  return UD_OK;
}

int ud_mpm_step_fwd(const ud_mpm_params* p, const ud_mpm_state* in, const int32_t* material, const float* h,
                    const float* action, ud_mpm_state* out, void* workspace, size_t workspace_bytes,
                    void* stream) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k)) return fail(UD_E_INVALID, "ud_mpm_step_fwd: invalid params");
  if (!state_ok(k, in) || !state_ok(k, out) || !material || !h || (k.n_prim > 0 && !action))
    return fail(UD_E_INVALID, "ud_mpm_step_fwd: null pointer");
  MpmWs ws;
  size_t need = mpm_carve(p, k, false, workspace, &ws);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255))
    return fail(UD_E_WORKSPACE, "ud_mpm_step_fwd: workspace too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;

  launch_sort(k, in->x, ws, nullptr, st);
  launch_gather_state(k, in, material, h, ws, ws.ps, st);
  launch_fk_fwd(k, in, action, out, ws, st);
  zero_async(ws.jrows, 4 * (size_t)k.B * k.S * 9, st);
  // the grids are zeroed in full once per call; between substeps only the 4x4x4 blocks P2G marked are re-zeroed
  zero_async(ws.blk_flag, 4 * (size_t)k.B * k.nbx * k.nby * k.nbz, st);
  zero_async(ws.blk_count, 8 * (size_t)k.S, st);
  const size_t BG = (size_t)k.B * k.G;
  zero_async(ws.grid_raw, 16 * BG * 2, st);
  if (ws.grid_fix) zero_async(ws.grid_fix, 32 * BG, st);
  auto vt_in = [&](int f) { return (g_svd_warm && (f % SVD_RESTART)) ? ws.vt_roll : nullptr; };
  const bool lists = p2g_lists_blocks();
  // Two grids, used alternately: substep f scatters into, updates and gathers from grid f % 2, and the grid launch of
  // substep f also re-zeroes the OTHER grid block by block from the list of substep f - 1 (whose G2P has finished), so
  // that P2G(f + 1) finds it empty: one grid-side launch per substep.
  for (int f = 0; f < k.S; ++f) {
    float4* gf = ws.grid_raw + BG * (f & 1);
    float4* gn = ws.grid_raw + BG * ((f + 1) & 1);
    launch_p2g(k, ws.ps, ws.ps, gf, in->mu, in->lamda, vt_in(f), ws.vt_roll, nullptr, f, ws, st);
    const bool clear = f >= 1 && f + 1 < k.S;
    launch_grid_fwd(k, gf, gf, ws.grid_fix, f, in, ws, st, clear ? gn : nullptr, f - 1, lists);
    launch_g2p(k, ws.ps, ws.ps, gf, f, ws, st);
  }
  launch_unsort_state(k, ws.ps, in->J, ws, out, st);
  cudaMemcpyAsync(out->friction, in->friction, 4 * (size_t)k.B, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(out->mu, in->mu, 4 * (size_t)k.B, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(out->lamda, in->lamda, 4 * (size_t)k.B, cudaMemcpyDeviceToDevice, st);
  if (cudaGetLastError() != cudaSuccess) return fail(UD_E_CUDA, "ud_mpm_step_fwd: launch failed");
  return UD_OK;
}

// Forward sweep that keeps every substep's start state, both grids, the SVDs and the active-cell lists in a
// bwd-layout workspace (the "tape"): the adjoint's recompute pass, and the whole of the taped forward.
static void mpm_prepare(const MpmConst& k, const ud_mpm_state* in, const int32_t* material, const float* h,
                        const float* action, ud_mpm_state* out, MpmWs& ws, float* ps0, cudaStream_t st) {
  launch_sort(k, in->x, ws, nullptr, st);
  launch_gather_state(k, in, material, h, ws, ps0, st);
  launch_fk_fwd(k, in, action, out, ws, st);
}
// Records substeps [f0, f1) into slots 0.. of the per-substep arrays (ws.sub0 must be f0).  start = state at f0
// (null: ws.ps slot 0 already holds it); vt0 = V^T entering f0 in the forward's buffer layout (windowed adjoint).
// need_end_state: the state after substep f1 - 1 is wanted (the taped forward un-sorts it into the output).  The adjoint
// never reads it -- it needs the substep's grids, not what G2P makes of them -- so its recompute pass skips that G2P.
static void mpm_record_range(const MpmConst& k, const ud_mpm_state* in, MpmWs& ws, int f0, int f1, const float* start,
                             const float* vt0, cudaStream_t st, bool need_end_state) {
  const size_t slot = (size_t)PS_NCOMP * k.N_pad, BG = (size_t)k.B * k.G;
  const int nf = f1 - f0;
  zero_async(ws.grid_raw, 16 * BG * 2, st);   // two raw grids, used alternately as in the forward (round 2 until its last
                                             // session kept one per substep: a 600 MB memset per call at S = 16, B = 32)
  if (ws.grid_fix) zero_async(ws.grid_fix, 32 * BG, st);   // consumed (re-zeroed) cell by cell by k_grid_fwd
  zero_async(ws.blk_flag, 4 * (size_t)k.B * k.nbx * k.nby * k.nbz, st);
  zero_async(ws.blk_count + 2 * f0, 8 * (size_t)nf, st);
  zero_async(ws.act_count + f0, 4 * (size_t)nf, st);
  auto sv = [&](int f) { return ws.svd_s + (size_t)SV_NCOMP * k.N_pad * (f - f0); };
  auto ps_rd = [&](int f) { return (f == f0 && start) ? start : ws.ps + slot * (f - f0); };
  auto ps_wr = [&](int f) { return ws.ps + slot * (f - f0); };
  const bool lists = p2g_lists_blocks();
  for (int f = f0; f < f1; ++f) {
    const bool warm = g_svd_warm && (f % SVD_RESTART);
    const bool from_ckpt = warm && f == f0 && vt0;
    const float* vt_in = !warm ? nullptr : (from_ckpt ? vt0 : sv(f - 1));
    float4* gf = ws.grid_raw + BG * (f & 1);
    float4* gn = ws.grid_raw + BG * ((f + 1) & 1);
    launch_p2g(k, ps_rd(f), ps_wr(f + 1), gf, in->mu, in->lamda, vt_in, nullptr, sv(f), f, ws, st, from_ckpt);
    // the raw {p, m} of the listed cells is kept in list order (ws.act_raw) for the grid adjoint; the other raw grid
    // (substep f - 1's, read for the last time by that substep's grid launch) is re-zeroed block by block here
    const bool clear = f > f0 && f + 1 < f1;
    launch_grid_fwd(k, gf, ws.grid_out + BG * (f - f0), ws.grid_fix, f, in, ws, st, clear ? gn : nullptr, f - 1, lists);
    if (f + 1 < f1 || need_end_state) launch_g2p(k, ps_rd(f), ps_wr(f + 1), ws.grid_out + BG * (f - f0), f, ws, st);
  }
}
static void mpm_record_pass(const MpmConst& k, const ud_mpm_state* in, const int32_t* material, const float* h,
                            const float* action, ud_mpm_state* out, MpmWs& ws, cudaStream_t st) {
  mpm_prepare(k, in, material, h, action, out, ws, ws.ps, st);
  ws.sub0 = 0;
  mpm_record_range(k, in, ws, 0, k.S, nullptr, nullptr, st, out != nullptr);
}

static void mpm_reverse_begin(const MpmConst& k, const ud_mpm_state* gout, MpmWs& ws, cudaStream_t st) {
  const size_t BG = (size_t)k.B * k.G;
  const size_t P = k.n_prim > 0 ? k.n_prim : 1;
  launch_gather_cot(k, gout, ws, st);
  zero_async(ws.g_fk_pos, 4 * (size_t)k.B * P * (k.S + 1) * 3, st);
  zero_async(ws.g_fk_rot, 4 * (size_t)k.B * P * (k.S + 1) * 4, st);
  zero_async(ws.g_fk_v, 4 * (size_t)k.B * P * k.S * 3, st);
  zero_async(ws.g_scal, 4 * (size_t)k.B * GS_STRIDE, st);
  zero_async(ws.g_prim_in, 4 * (size_t)k.B * P * 16, st);
  zero_async(ws.g_act, 4 * (size_t)k.B * P * 6, st);
  zero_async(ws.ggrid, 16 * BG * 2, st);
}
// Reverses substeps f1-1 .. f0 recorded in slots 0.. (ws.sub0 == f0).  start = state at f0 when it is not in slot 0.
// Two cotangent grids, used alternately: G2P^T of substep f scatters into slot f & 1, the grid adjoint reverses it in
// place and, in the same launch, re-zeroes the OTHER slot from the block list of substep f + 1 (whose P2G^T has
// finished), so that G2P^T(f - 1) finds it empty -- G2P^T(f + 1) scattered into the blocks P2G(f + 1) marked (+ face
// cells through clamped indices, when that substep's shell flag is up).
// first_of_call: both slots are still as mpm_reverse_begin zeroed them.  Otherwise they hold cotangents of substeps
// whose block lists belong to a window that is gone: both are re-zeroed in full.
static void mpm_reverse_range(const MpmConst& k, const ud_mpm_state* in, MpmWs& ws, int f0, int f1, const float* start,
                              bool first_of_call, cudaStream_t st) {
  const size_t slot = (size_t)PS_NCOMP * k.N_pad, BG = (size_t)k.B * k.G;
  if (!first_of_call) zero_async(ws.ggrid, 16 * BG * 2, st);
  for (int f = f1 - 1; f >= f0; --f) {
    const float* s_in = (f == f0 && start) ? start : ws.ps + slot * (f - f0);
    float4* gg = ws.ggrid + BG * (f & 1);
    float4* other = ws.ggrid + BG * ((f + 1) & 1);
    launch_g2p_bwd(k, s_in, ws.grid_out + BG * (f - f0), ws, st, gg);
    launch_grid_bwd(k, f, in, ws, st, gg, f < f1 - 1 ? other : nullptr, f + 1);
    launch_p2g_bwd(k, s_in, ws.svd_s + (size_t)SV_NCOMP * k.N_pad * (f - f0), in->mu, in->lamda, f == 0, ws, st, gg);
  }
}
static void mpm_reverse_end(const MpmConst& k, const ud_mpm_state* in, const float* action, const ud_mpm_state* gout,
                            ud_mpm_state* gin, float* gaction, MpmWs& ws, cudaStream_t st) {
  launch_fk_bwd(k, in, action, gout, ws, st);
  launch_finish_bwd(k, in, gout, gin, action, gaction, ws, st);
}
static void mpm_reverse_pass(const MpmConst& k, const ud_mpm_state* in, const float* action, const ud_mpm_state* gout,
                             ud_mpm_state* gin, float* gaction, MpmWs& ws, cudaStream_t st) {
  ws.sub0 = 0;
  mpm_reverse_begin(k, gout, ws, st);
  mpm_reverse_range(k, in, ws, 0, k.S, nullptr, true, st);
  mpm_reverse_end(k, in, action, gout, gin, gaction, ws, st);
}

int ud_mpm_step_bwd(const ud_mpm_params* p, const ud_mpm_state* in, const int32_t* material, const float* h,
                    const float* action, const ud_mpm_state* gout, ud_mpm_state* gin, float* gaction,
                    void* workspace, size_t workspace_bytes, void* stream) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k)) return fail(UD_E_INVALID, "ud_mpm_step_bwd: invalid params");
  if (!state_ok(k, in) || !gout || !gin || !material || !h || (k.n_prim > 0 && !action))
    return fail(UD_E_INVALID, "ud_mpm_step_bwd: null pointer");
  MpmWs ws;
  size_t need = mpm_carve(p, k, true, workspace, &ws);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255))
    return fail(UD_E_WORKSPACE, "ud_mpm_step_bwd: workspace too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  mpm_record_pass(k, in, material, h, action, nullptr, ws, st);   // checkpoint = the step input
  mpm_reverse_pass(k, in, action, gout, gin, gaction, ws, st);
  if (cudaGetLastError() != cudaSuccess) return fail(UD_E_CUDA, "ud_mpm_step_bwd: launch failed");
  return UD_OK;
}

size_t ud_mpm_bwd_windowed_workspace_bytes(const ud_mpm_params* p, int32_t window) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k) || window < 1) return 0;
  if (window >= k.S) return ud_mpm_bwd_workspace_bytes(p);
  return mpm_carve(p, k, true, nullptr, nullptr, window);
}

int ud_mpm_step_bwd_windowed(const ud_mpm_params* p, const ud_mpm_state* in, const int32_t* material, const float* h,
                             const float* action, const ud_mpm_state* gout, ud_mpm_state* gin, float* gaction,
                             int32_t window, void* workspace, size_t workspace_bytes, void* stream) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k) || window < 1) return fail(UD_E_INVALID, "ud_mpm_step_bwd_windowed: invalid params");
  if (window >= k.S)
    return ud_mpm_step_bwd(p, in, material, h, action, gout, gin, gaction, workspace, workspace_bytes, stream);
  if (!state_ok(k, in) || !gout || !gin || !material || !h || (k.n_prim > 0 && !action))
    return fail(UD_E_INVALID, "ud_mpm_step_bwd_windowed: null pointer");
  MpmWs ws;
  size_t need = mpm_carve(p, k, true, workspace, &ws, window);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255))
    return fail(UD_E_WORKSPACE, "ud_mpm_step_bwd_windowed: workspace too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int W = window, S = k.S, n_win = (S + W - 1) / W;
  const size_t slot = (size_t)PS_NCOMP * k.N_pad, vslot = (size_t)VT_NCOMP * k.N_pad, BG = (size_t)k.B * k.G;
  // ---- checkpoint pass: the in-place forward from the step input, keeping the state entering every W-th substep
  mpm_prepare(k, in, material, h, action, nullptr, ws, ws.ckpt_ps, st);
  cudaMemcpyAsync(ws.run_ps, ws.ckpt_ps, 4 * slot, cudaMemcpyDeviceToDevice, st);
  zero_async(ws.blk_flag, 4 * (size_t)k.B * k.nbx * k.nby * k.nbz, st);
  zero_async(ws.blk_count, 8 * (size_t)S, st);
  zero_async(ws.run_grid, 16 * BG * 2, st);
  if (ws.grid_fix) zero_async(ws.grid_fix, 32 * BG, st);
  {
    const bool lists = p2g_lists_blocks();
    const int last = (n_win - 1) * W;   // nothing after the last checkpoint is needed from this pass
    for (int f = 0; f < last; ++f) {
      float4* gf = ws.run_grid + BG * (f & 1);
      float4* gn = ws.run_grid + BG * ((f + 1) & 1);
      const float* vt_in = (g_svd_warm && (f % SVD_RESTART)) ? ws.run_vt : nullptr;
      launch_p2g(k, ws.run_ps, ws.run_ps, gf, in->mu, in->lamda, vt_in, ws.run_vt, nullptr, f, ws, st);
      launch_grid_fwd(k, gf, gf, ws.grid_fix, f, in, ws, st, (f >= 1 && f + 1 < last) ? gn : nullptr, f - 1, lists);
      launch_g2p(k, ws.run_ps, ws.run_ps, gf, f, ws, st);
      if ((f + 1) % W == 0) {
        const int w = (f + 1) / W;
        cudaMemcpyAsync(ws.ckpt_ps + slot * w, ws.run_ps, 4 * slot, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(ws.ckpt_vt + vslot * w, ws.run_vt, 4 * vslot, cudaMemcpyDeviceToDevice, st);
      }
    }
  }
  // ---- windows, last first: recompute W substeps from their checkpoint, reverse them
  mpm_reverse_begin(k, gout, ws, st);
  for (int w = n_win - 1; w >= 0; --w) {
    const int f0 = w * W, f1 = f0 + W < S ? f0 + W : S;
    ws.sub0 = f0;
    mpm_record_range(k, in, ws, f0, f1, ws.ckpt_ps + slot * w, w ? ws.ckpt_vt + vslot * w : nullptr, st, false);
    mpm_reverse_range(k, in, ws, f0, f1, ws.ckpt_ps + slot * w, w == n_win - 1, st);
  }
  mpm_reverse_end(k, in, action, gout, gin, gaction, ws, st);
  if (cudaGetLastError() != cudaSuccess) return fail(UD_E_CUDA, "ud_mpm_step_bwd_windowed: launch failed");
  return UD_OK;
}

size_t ud_mpm_tape_bytes(const ud_mpm_params* p) { return ud_mpm_bwd_workspace_bytes(p); }

int ud_mpm_step_fwd_taped(const ud_mpm_params* p, const ud_mpm_state* in, const int32_t* material, const float* h,
                          const float* action, ud_mpm_state* out, void* tape, size_t tape_bytes, void* stream) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k)) return fail(UD_E_INVALID, "ud_mpm_step_fwd_taped: invalid params");
  if (!state_ok(k, in) || !state_ok(k, out) || !material || !h || (k.n_prim > 0 && !action))
    return fail(UD_E_INVALID, "ud_mpm_step_fwd_taped: null pointer");
  MpmWs ws;
  size_t need = mpm_carve(p, k, true, tape, &ws);
  if (!tape || tape_bytes < need || ((uintptr_t)tape & 255))
    return fail(UD_E_WORKSPACE, "ud_mpm_step_fwd_taped: tape too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  mpm_record_pass(k, in, material, h, action, out, ws, st);
  launch_unsort_state(k, ws.ps + (size_t)PS_NCOMP * k.N_pad * k.S, in->J, ws, out, st);
  cudaMemcpyAsync(out->friction, in->friction, 4 * (size_t)k.B, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(out->mu, in->mu, 4 * (size_t)k.B, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(out->lamda, in->lamda, 4 * (size_t)k.B, cudaMemcpyDeviceToDevice, st);
  if (cudaGetLastError() != cudaSuccess) return fail(UD_E_CUDA, "ud_mpm_step_fwd_taped: launch failed");
  return UD_OK;
}

int ud_mpm_step_bwd_taped(const ud_mpm_params* p, const ud_mpm_state* in, const float* action,
                          const ud_mpm_state* gout, ud_mpm_state* gin, float* gaction, void* tape, size_t tape_bytes,
                          void* stream) {
  MpmConst k;
  if (!mpm_fold_constants(p, &k)) return fail(UD_E_INVALID, "ud_mpm_step_bwd_taped: invalid params");
  if (!state_ok(k, in) || !gout || !gin || (k.n_prim > 0 && !action))
    return fail(UD_E_INVALID, "ud_mpm_step_bwd_taped: null pointer");
  MpmWs ws;
  size_t need = mpm_carve(p, k, true, tape, &ws);
  if (!tape || tape_bytes < need || ((uintptr_t)tape & 255))
    return fail(UD_E_WORKSPACE, "ud_mpm_step_bwd_taped: tape too small or misaligned");
  mpm_reverse_pass(k, in, action, gout, gin, gaction, ws, (cudaStream_t)stream);
  if (cudaGetLastError() != cudaSuccess) return fail(UD_E_CUDA, "ud_mpm_step_bwd_taped: launch failed");
  return UD_OK;
}

}  // extern "C"
