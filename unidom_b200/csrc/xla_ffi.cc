// XLA-FFI adapter over the C ABI (include/unidom_b200.h) for jax.ffi.ffi_call + jax.custom_vjp.
//
// Compiled only where "xla/ffi/api/ffi.h" is on the include path: unidom_b200/build.py adds jax.ffi.include_dir() when
// jax is importable and builds libunidom_b200_xla.so.  Neither jax nor the XLA headers exist in THIS image, so here the
// file is compiled against tests/xla_stub (a test-only stand-in for the API surface used below) and its handlers are
// CALLED through that stub by tests/test_xla_ffi.py -- buffer order, attribute decoding, workspace plumbing and error
// mapping are verified; ABI compatibility with a real XLA is not.  The file contains no arithmetic:
// every handler forwards raw device pointers to the ud_* entry points.
//
// Buffer order = the reference's pytree flattening order (NamedTuple field order):
//   MPMState       core/engine/mpm_simulator.py:13-24   x v C F J [cur_step] primitives... [key] friction mu lamda
//   PrimitiveState core/engine/primitives/primitives.py:9-23
//   ClothState     core/engine/cloth_simulator.py:13-23
// The Python wrapper passes only the float leaves, in the order of ud_mpm_state / ud_cloth_state.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define UD_HAVE_XLA_FFI 1
#endif
#endif

#ifdef UD_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/unidom_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

constexpr int kStateLeaves = 8;  // x v C F J friction mu lamda
constexpr int kPrimLeaves = 9;   // size friction softness position rotation v w action_buffer action_scale

template <class Getter>
void fill_mpm_state(ud_mpm_state* s, int n_prim, int first, Getter get) {
  std::memset(s, 0, sizeof(*s));
  float** f = reinterpret_cast<float**>(s);
  for (int i = 0; i < kStateLeaves; ++i) f[i] = static_cast<float*>(get(first + i));
  for (int q = 0; q < n_prim; ++q) {
    float** pf = reinterpret_cast<float**>(&s->prim[q]);
    for (int i = 0; i < kPrimLeaves; ++i) pf[i] = static_cast<float*>(get(first + kStateLeaves + q * kPrimLeaves + i));
  }
}

ffi::Error status(int rc, const char* what) {
  if (rc == UD_OK) return ffi::Error::Success();
  return ffi::Error(rc == UD_E_CUDA ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument,
                    std::string(what) + ": " + ud_last_error());
}

ud_mpm_params mpm_params(int32_t B, int32_t n, int32_t steps, int32_t rx, int32_t ry, int32_t rz, int32_t n_grid,
                         double dt, double p_rho, double gx, double gy, double gz, int32_t n_prim, int32_t sdf_kind,
                         int32_t pos_control, int32_t p2g_mode) {
  ud_mpm_params p;
  std::memset(&p, 0, sizeof(p));
  p.num_envs = B; p.n_particles = n; p.steps = steps;
  p.res[0] = rx; p.res[1] = ry; p.res[2] = rz; p.n_grid = n_grid;
  p.dt = dt; p.dx = 1.0 / n_grid; p.inv_dx = (double)n_grid;
  p.p_vol = (p.dx * 0.5) * (p.dx * 0.5); p.p_mass = p.p_vol * p_rho;
  p.gravity[0] = gx; p.gravity[1] = gy; p.gravity[2] = gz;
  p.n_primitive = n_prim; p.sdf_kind = sdf_kind; p.use_position_control = pos_control; p.p2g_mode = p2g_mode;
  return p;
}

// args: material(i32[n]) h(f32[n]) action(f32[B,6P]) then the input leaves; rets: output leaves, workspace(u8[...])
ffi::Error MpmFwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t steps, int32_t rx,
                  int32_t ry, int32_t rz, int32_t n_grid, double dt, double p_rho, double gx, double gy, double gz,
                  int32_t n_prim, int32_t sdf_kind, int32_t pos_control, int32_t p2g_mode) {
  auto x = args.get<ffi::AnyBuffer>(3).value();
  auto dims = x.dimensions();
  ud_mpm_params p = mpm_params((int32_t)dims[0], (int32_t)dims[1], steps, rx, ry, rz, n_grid, dt, p_rho, gx, gy, gz,
                               n_prim, sdf_kind, pos_control, p2g_mode);
  ud_mpm_state in, out;
  fill_mpm_state(&in, n_prim, 3, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_mpm_state(&out, n_prim, 0, [&](int i) { return rets.get<ffi::AnyBuffer>(i).value()->untyped_data(); });
  auto ws = rets.get<ffi::AnyBuffer>(rets.size() - 1).value();
  int rc = ud_mpm_step_fwd(&p, &in, static_cast<const int32_t*>(args.get<ffi::AnyBuffer>(0).value().untyped_data()),
                           static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()),
                           static_cast<const float*>(args.get<ffi::AnyBuffer>(2).value().untyped_data()), &out,
                           ws->untyped_data(), ws->size_bytes(), stream);
  return status(rc, "ud_mpm_step_fwd");
}

// args: material h action, input leaves, output-cotangent leaves; rets: input-cotangent leaves, gaction, workspace
ffi::Error MpmBwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t steps, int32_t rx,
                  int32_t ry, int32_t rz, int32_t n_grid, double dt, double p_rho, double gx, double gy, double gz,
                  int32_t n_prim, int32_t sdf_kind, int32_t pos_control, int32_t p2g_mode) {
  auto x = args.get<ffi::AnyBuffer>(3).value();
  auto dims = x.dimensions();
  ud_mpm_params p = mpm_params((int32_t)dims[0], (int32_t)dims[1], steps, rx, ry, rz, n_grid, dt, p_rho, gx, gy, gz,
                               n_prim, sdf_kind, pos_control, p2g_mode);
  const int L = kStateLeaves + n_prim * kPrimLeaves;
  ud_mpm_state in, gout, gin;
  fill_mpm_state(&in, n_prim, 3, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_mpm_state(&gout, n_prim, 3 + L, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_mpm_state(&gin, n_prim, 0, [&](int i) { return rets.get<ffi::AnyBuffer>(i).value()->untyped_data(); });
  auto gaction = rets.get<ffi::AnyBuffer>(L).value();
  auto ws = rets.get<ffi::AnyBuffer>(L + 1).value();
  int rc = ud_mpm_step_bwd(&p, &in, static_cast<const int32_t*>(args.get<ffi::AnyBuffer>(0).value().untyped_data()),
                           static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()),
                           static_cast<const float*>(args.get<ffi::AnyBuffer>(2).value().untyped_data()), &gout, &gin,
                           static_cast<float*>(gaction->untyped_data()), ws->untyped_data(), ws->size_bytes(), stream);
  return status(rc, "ud_mpm_step_bwd");
}

ud_cloth_params cloth_params(int32_t B, int32_t P, int32_t N, double dt, double gravity, double damping, double max_v,
                             double small_num, double mask_sum, int32_t stiff_float) {
  ud_cloth_params p;
  std::memset(&p, 0, sizeof(p));
  p.num_envs = B; p.n_nodes = P; p.N = N; p.substeps = 50;
  p.dt = dt; p.gravity = gravity; p.damping = damping; p.max_v = max_v; p.small_num = small_num;
  p.cell_size = 1.0 / N; p.mask_sum = mask_sum; p.stiffness_is_float = stiff_float;
  return p;
}

template <class Getter>
void fill_cloth_state(ud_cloth_state* s, int first, Getter get) {
  float** f = reinterpret_cast<float**>(s);
  for (int i = 0; i < 8; ++i) f[i] = static_cast<float*>(get(first + i));
}

// args: nbr(i32[P,8]) L0(f32[P,8]) action(f32[B,8]) then the 8 input leaves; rets: 8 output leaves
ffi::Error ClothFwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t N, double dt,
                    double gravity, double damping, double max_v, double small_num, double mask_sum,
                    int32_t stiff_float) {
  auto dims = args.get<ffi::AnyBuffer>(3).value().dimensions();
  ud_cloth_params p = cloth_params((int32_t)dims[0], (int32_t)dims[1], N, dt, gravity, damping, max_v, small_num,
                                   mask_sum, stiff_float);
  ud_cloth_state in, out;
  fill_cloth_state(&in, 3, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_cloth_state(&out, 0, [&](int i) { return rets.get<ffi::AnyBuffer>(i).value()->untyped_data(); });
  int rc = ud_cloth_step_fwd(&p, &in, static_cast<const int32_t*>(args.get<ffi::AnyBuffer>(0).value().untyped_data()),
                             static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()),
                             static_cast<const float*>(args.get<ffi::AnyBuffer>(2).value().untyped_data()), &out,
                             nullptr, 0, stream);
  return status(rc, "ud_cloth_step_fwd");
}

// args: nbr L0 action, 8 input leaves, 8 output cotangents; rets: 8 input cotangents, gaction, workspace
ffi::Error ClothBwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t N, double dt,
                    double gravity, double damping, double max_v, double small_num, double mask_sum,
                    int32_t stiff_float) {
  auto dims = args.get<ffi::AnyBuffer>(3).value().dimensions();
  ud_cloth_params p = cloth_params((int32_t)dims[0], (int32_t)dims[1], N, dt, gravity, damping, max_v, small_num,
                                   mask_sum, stiff_float);
  ud_cloth_state in, gout, gin;
  fill_cloth_state(&in, 3, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_cloth_state(&gout, 11, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_cloth_state(&gin, 0, [&](int i) { return rets.get<ffi::AnyBuffer>(i).value()->untyped_data(); });
  auto gaction = rets.get<ffi::AnyBuffer>(8).value();
  auto ws = rets.get<ffi::AnyBuffer>(9).value();
  int rc = ud_cloth_step_bwd(&p, &in, static_cast<const int32_t*>(args.get<ffi::AnyBuffer>(0).value().untyped_data()),
                             static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()),
                             static_cast<const float*>(args.get<ffi::AnyBuffer>(2).value().untyped_data()), &gout, &gin,
                             static_cast<float*>(gaction->untyped_data()), ws->untyped_data(), ws->size_bytes(), stream);
  return status(rc, "ud_cloth_step_bwd");
}

// ---- taped MPM pair: the tape is a ret of the forward and an arg of the backward
// args: material h action, input leaves; rets: output leaves, tape(u8[ud_mpm_tape_bytes])
ffi::Error MpmFwdTaped(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t steps, int32_t rx,
                       int32_t ry, int32_t rz, int32_t n_grid, double dt, double p_rho, double gx, double gy, double gz,
                       int32_t n_prim, int32_t sdf_kind, int32_t pos_control, int32_t p2g_mode) {
  auto dims = args.get<ffi::AnyBuffer>(3).value().dimensions();
  ud_mpm_params p = mpm_params((int32_t)dims[0], (int32_t)dims[1], steps, rx, ry, rz, n_grid, dt, p_rho, gx, gy, gz,
                               n_prim, sdf_kind, pos_control, p2g_mode);
  ud_mpm_state in, out;
  fill_mpm_state(&in, n_prim, 3, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_mpm_state(&out, n_prim, 0, [&](int i) { return rets.get<ffi::AnyBuffer>(i).value()->untyped_data(); });
  auto tape = rets.get<ffi::AnyBuffer>(rets.size() - 1).value();
  int rc = ud_mpm_step_fwd_taped(&p, &in, static_cast<const int32_t*>(args.get<ffi::AnyBuffer>(0).value().untyped_data()),
                                 static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()),
                                 static_cast<const float*>(args.get<ffi::AnyBuffer>(2).value().untyped_data()), &out,
                                 tape->untyped_data(), tape->size_bytes(), stream);
  return status(rc, "ud_mpm_step_fwd_taped");
}
// args: action, input leaves, output-cotangent leaves, tape; rets: input-cotangent leaves, gaction
ffi::Error MpmBwdTaped(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t steps, int32_t rx,
                       int32_t ry, int32_t rz, int32_t n_grid, double dt, double p_rho, double gx, double gy, double gz,
                       int32_t n_prim, int32_t sdf_kind, int32_t pos_control, int32_t p2g_mode) {
  auto dims = args.get<ffi::AnyBuffer>(1).value().dimensions();
  ud_mpm_params p = mpm_params((int32_t)dims[0], (int32_t)dims[1], steps, rx, ry, rz, n_grid, dt, p_rho, gx, gy, gz,
                               n_prim, sdf_kind, pos_control, p2g_mode);
  const int L = kStateLeaves + n_prim * kPrimLeaves;
  ud_mpm_state in, gout, gin;
  fill_mpm_state(&in, n_prim, 1, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_mpm_state(&gout, n_prim, 1 + L, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_mpm_state(&gin, n_prim, 0, [&](int i) { return rets.get<ffi::AnyBuffer>(i).value()->untyped_data(); });
  auto tape = args.get<ffi::AnyBuffer>(1 + 2 * L).value();
  auto gaction = rets.get<ffi::AnyBuffer>(L).value();
  int rc = ud_mpm_step_bwd_taped(&p, &in, static_cast<const float*>(args.get<ffi::AnyBuffer>(0).value().untyped_data()), &gout,
                                 &gin, static_cast<float*>(gaction->untyped_data()), tape.untyped_data(), tape.size_bytes(),
                                 stream);
  return status(rc, "ud_mpm_step_bwd_taped");
}

// ---- fused cloth env step: T sub-actions per call (cloth_env.py:211)
// args: nbr L0 actions(f32[B,T,8]) then the 8 input leaves; rets: 8 output leaves, ckpt(u8[...])
ffi::Error ClothMultiFwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t N, double dt,
                         double gravity, double damping, double max_v, double small_num, double mask_sum,
                         int32_t stiff_float) {
  auto dims = args.get<ffi::AnyBuffer>(3).value().dimensions();
  const int32_t T = (int32_t)args.get<ffi::AnyBuffer>(2).value().dimensions()[1];
  ud_cloth_params p = cloth_params((int32_t)dims[0], (int32_t)dims[1], N, dt, gravity, damping, max_v, small_num,
                                   mask_sum, stiff_float);
  ud_cloth_state in, out;
  fill_cloth_state(&in, 3, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_cloth_state(&out, 0, [&](int i) { return rets.get<ffi::AnyBuffer>(i).value()->untyped_data(); });
  auto ck = rets.get<ffi::AnyBuffer>(8).value();
  int rc = ud_cloth_multi_step_fwd(&p, &in, static_cast<const int32_t*>(args.get<ffi::AnyBuffer>(0).value().untyped_data()),
                                   static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()),
                                   static_cast<const float*>(args.get<ffi::AnyBuffer>(2).value().untyped_data()), T, &out,
                                   ck->untyped_data(), ck->size_bytes(), stream);
  return status(rc, "ud_cloth_multi_step_fwd");
}
// args: nbr L0 actions, 8 input leaves, ckpt, 8 output cotangents; rets: 8 input cotangents, gactions, workspace
ffi::Error ClothMultiBwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t N, double dt,
                         double gravity, double damping, double max_v, double small_num, double mask_sum,
                         int32_t stiff_float) {
  auto dims = args.get<ffi::AnyBuffer>(3).value().dimensions();
  const int32_t T = (int32_t)args.get<ffi::AnyBuffer>(2).value().dimensions()[1];
  ud_cloth_params p = cloth_params((int32_t)dims[0], (int32_t)dims[1], N, dt, gravity, damping, max_v, small_num,
                                   mask_sum, stiff_float);
  ud_cloth_state in, gout, gin;
  fill_cloth_state(&in, 3, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_cloth_state(&gout, 12, [&](int i) { return args.get<ffi::AnyBuffer>(i).value().untyped_data(); });
  fill_cloth_state(&gin, 0, [&](int i) { return rets.get<ffi::AnyBuffer>(i).value()->untyped_data(); });
  auto gact = rets.get<ffi::AnyBuffer>(8).value();
  auto ws = rets.get<ffi::AnyBuffer>(9).value();
  int rc = ud_cloth_multi_step_bwd(&p, &in, static_cast<const int32_t*>(args.get<ffi::AnyBuffer>(0).value().untyped_data()),
                                   static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()),
                                   static_cast<const float*>(args.get<ffi::AnyBuffer>(2).value().untyped_data()), T,
                                   args.get<ffi::AnyBuffer>(11).value().untyped_data(), &gout, &gin,
                                   static_cast<float*>(gact->untyped_data()), ws->untyped_data(), ws->size_bytes(), stream);
  return status(rc, "ud_cloth_multi_step_bwd");
}

// ---- rewards (core/utils/util.py:138-159).  chamfer: args x(f32[B,P,3]) y(f32[Q,3]); rets out(f32[B]) residuals(u8)
ffi::Error ChamferFwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  auto x = args.get<ffi::AnyBuffer>(0).value();
  auto y = args.get<ffi::AnyBuffer>(1).value();
  auto res = rets.get<ffi::AnyBuffer>(1).value();
  int rc = ud_chamfer_fwd(static_cast<const float*>(x.untyped_data()), static_cast<const float*>(y.untyped_data()),
                          (int32_t)x.dimensions()[0], (int32_t)x.dimensions()[1], (int32_t)y.dimensions()[0],
                          static_cast<float*>(rets.get<ffi::AnyBuffer>(0).value()->untyped_data()), res->untyped_data(),
                          res->size_bytes(), stream);
  return status(rc, "ud_chamfer_fwd");
}
// args x y gout(f32[B]) residuals; rets gx(f32[B,P,3])
ffi::Error ChamferBwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  auto x = args.get<ffi::AnyBuffer>(0).value();
  auto y = args.get<ffi::AnyBuffer>(1).value();
  auto res = args.get<ffi::AnyBuffer>(3).value();
  int rc = ud_chamfer_bwd(static_cast<const float*>(x.untyped_data()), static_cast<const float*>(y.untyped_data()),
                          (int32_t)x.dimensions()[0], (int32_t)x.dimensions()[1], (int32_t)y.dimensions()[0],
                          static_cast<const float*>(args.get<ffi::AnyBuffer>(2).value().untyped_data()), res.untyped_data(),
                          res.size_bytes(), static_cast<float*>(rets.get<ffi::AnyBuffer>(0).value()->untyped_data()), stream);
  return status(rc, "ud_chamfer_bwd");
}
// l2: args x(f32[B,P,3]) y(f32[P,3]); rets out(f32[B])
ffi::Error L2Fwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  auto x = args.get<ffi::AnyBuffer>(0).value();
  int rc = ud_l2_fwd(static_cast<const float*>(x.untyped_data()),
                     static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()), (int32_t)x.dimensions()[0],
                     (int32_t)x.dimensions()[1], static_cast<float*>(rets.get<ffi::AnyBuffer>(0).value()->untyped_data()), stream);
  return status(rc, "ud_l2_fwd");
}
// args x y gout(f32[B]); rets gx
ffi::Error L2Bwd(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  auto x = args.get<ffi::AnyBuffer>(0).value();
  int rc = ud_l2_bwd(static_cast<const float*>(x.untyped_data()),
                     static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()), (int32_t)x.dimensions()[0],
                     (int32_t)x.dimensions()[1], static_cast<const float*>(args.get<ffi::AnyBuffer>(2).value().untyped_data()),
                     static_cast<float*>(rets.get<ffi::AnyBuffer>(0).value()->untyped_data()), stream);
  return status(rc, "ud_l2_bwd");
}

// ---- APG update on the flat gradient (algorithms/apg/apg.py:233-240, 260-267).  XLA buffers are immutable inputs /
// fresh outputs: the handlers copy input -> output on the stream and update the output in place.
// scrub_clip: args grad(f32[n]); rets grad_out(f32[n]) sumsq(f32[1]); attr max_grad_norm
ffi::Error ApgScrubClip(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, double max_grad_norm) {
  auto g = args.get<ffi::AnyBuffer>(0).value();
  auto go = rets.get<ffi::AnyBuffer>(0).value();
  cudaMemcpyAsync(go->untyped_data(), g.untyped_data(), g.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  int rc = ud_apg_scrub_clip(static_cast<float*>(go->untyped_data()), (int64_t)g.element_count(), (float)max_grad_norm,
                             static_cast<float*>(rets.get<ffi::AnyBuffer>(1).value()->untyped_data()), stream);
  return status(rc, "ud_apg_scrub_clip");
}
// adam: args params grad m v (f32[n] each); rets params' m' v'; attrs world_size lr b1 b2 eps t
ffi::Error AdamStep(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t world_size, double lr,
                    double b1, double b2, double eps, int32_t t) {
  auto p = args.get<ffi::AnyBuffer>(0).value();
  const size_t bytes = p.size_bytes();
  void* outs[3];
  const int src[3] = {0, 2, 3};
  for (int i = 0; i < 3; ++i) {
    outs[i] = rets.get<ffi::AnyBuffer>(i).value()->untyped_data();
    cudaMemcpyAsync(outs[i], args.get<ffi::AnyBuffer>(src[i]).value().untyped_data(), bytes, cudaMemcpyDeviceToDevice, stream);
  }
  int rc = ud_adam_step(static_cast<float*>(outs[0]), static_cast<const float*>(args.get<ffi::AnyBuffer>(1).value().untyped_data()),
                        static_cast<float*>(outs[1]), static_cast<float*>(outs[2]), (int64_t)p.element_count(), world_size, lr,
                        b1, b2, eps, t, stream);
  return status(rc, "ud_adam_step");
}

#define UD_PLAIN_BIND() ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().RemainingArgs().RemainingRets()
#define UD_MPM_BIND()                                                                                         \
  ffi::Ffi::Bind()                                                                                            \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                               \
      .RemainingArgs()                                                                                        \
      .RemainingRets()                                                                                        \
      .Attr<int32_t>("steps").Attr<int32_t>("res_x").Attr<int32_t>("res_y").Attr<int32_t>("res_z")            \
      .Attr<int32_t>("n_grid").Attr<double>("dt").Attr<double>("p_rho")                                       \
      .Attr<double>("gravity_x").Attr<double>("gravity_y").Attr<double>("gravity_z")                          \
      .Attr<int32_t>("n_primitive").Attr<int32_t>("sdf_kind").Attr<int32_t>("use_position_control")           \
      .Attr<int32_t>("p2g_mode")
#define UD_CLOTH_BIND()                                                                                       \
  ffi::Ffi::Bind()                                                                                            \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                               \
      .RemainingArgs()                                                                                        \
      .RemainingRets()                                                                                        \
      .Attr<int32_t>("N").Attr<double>("dt").Attr<double>("gravity").Attr<double>("damping")                  \
      .Attr<double>("max_v").Attr<double>("small_num").Attr<double>("mask_sum").Attr<int32_t>("stiffness_is_float")

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_mpm_step_fwd, MpmFwd, UD_MPM_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_mpm_step_bwd, MpmBwd, UD_MPM_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_cloth_step_fwd, ClothFwd, UD_CLOTH_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_cloth_step_bwd, ClothBwd, UD_CLOTH_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_mpm_step_fwd_taped, MpmFwdTaped, UD_MPM_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_mpm_step_bwd_taped, MpmBwdTaped, UD_MPM_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_cloth_multi_step_fwd, ClothMultiFwd, UD_CLOTH_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_cloth_multi_step_bwd, ClothMultiBwd, UD_CLOTH_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_chamfer_fwd, ChamferFwd, UD_PLAIN_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_chamfer_bwd, ChamferBwd, UD_PLAIN_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_l2_fwd, L2Fwd, UD_PLAIN_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_l2_bwd, L2Bwd, UD_PLAIN_BIND());
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_apg_scrub_clip, ApgScrubClip, UD_PLAIN_BIND().Attr<double>("max_grad_norm"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(ud_xla_adam_step, AdamStep,
                              UD_PLAIN_BIND().Attr<int32_t>("world_size").Attr<double>("lr").Attr<double>("b1")
                                  .Attr<double>("b2").Attr<double>("eps").Attr<int32_t>("t"));

#ifdef XLA_FFI_STUB
// Plain C entry point of the TEST stub (tests/xla_stub): builds a call frame from flat arrays and invokes `handler`.
// buffers: n_args + n_rets entries of {data, rank, dims[rank]...} described by ptrs / ranks / dims (concatenated) /
// elem_bytes; attributes by name with an int64 or a double value.  Returns the handler's error code; err receives the
// message (truncated to err_cap).
extern "C" int ud_stub_call(int (*handler)(::xla::ffi::StubFrame*), void* stream, int n_args, int n_rets, void* const* ptrs,
                            const int* ranks, const int64_t* dims, const int* elem_bytes, int n_attrs,
                            const char* const* attr_names, const int* attr_is_double, const int64_t* attr_i,
                            const double* attr_d, char* err, int err_cap) {
  ::xla::ffi::StubFrame f;
  f.stream = stream;
  size_t off = 0;
  for (int i = 0; i < n_args + n_rets; ++i) {
    ::xla::ffi::AnyBuffer b(ptrs[i], dims + off, (size_t)ranks[i], (size_t)elem_bytes[i]);
    (i < n_args ? f.args : f.rets).push_back(b);
    off += (size_t)ranks[i];
  }
  for (int i = 0; i < n_attrs; ++i) f.attrs.push_back({attr_names[i], attr_is_double[i] != 0, attr_i[i], attr_d[i]});
  const int rc = handler(&f);
  if (err && err_cap > 0) {
    std::strncpy(err, f.error.c_str(), (size_t)err_cap - 1);
    err[err_cap - 1] = 0;
  }
  return rc;
}
#endif

#endif  // UD_HAVE_XLA_FFI
