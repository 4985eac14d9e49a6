// XLA-FFI bindings of the C ABI (include/jaxib_b200.h) -- the binding a jax_ib maintainer adds so that the
// reference's host side stays Python/JAX: one custom call per pluggable stage of
// equations.py:194-208 plus the fused step.
//
// jaxlib's headers (xla/ffi/api/ffi.h) are NOT present in this image, so this file is compile-guarded and is not
// part of libjaxib_b200.so.  tests/test_cpu_host.py type-checks it against an API-shaped stand-in header
// (tests/xla_ffi_stub/, g++ -fsyntax-only); it has never run under XLA.  It is pure argument plumbing over the
// tested C entry points: XLA owns the buffers (inputs immutable, outputs pre-allocated, in-place only through
// input_output_aliases) and hands us its CUDA stream and a scratch allocator.
//
//   build (where jaxlib is installed):
//     nvcc -std=c++17 -shared -Xcompiler -fPIC -I$(python -c "import jaxlib, os; print(os.path.join(os.path.dirname(jaxlib.__file__), 'include'))") ...
//            xla_ffi_shim.cc -L../_native -ljaxib_b200 -o libjaxib_b200_ffi.so
//   register (Python):
//     jax.ffi.register_ffi_target("jaxib_step", jax.ffi.pycapsule(lib.JaxibStep), platform="CUDA")
//
// Every handler takes the grid as ATTRIBUTES (an ffi::Dictionary, the fields of struct jaxib_grid under their
// own names; missing ones keep the struct's zero / default) -- wall values, body force, advection scheme and
// boundary family included.  Plans are cached on EVERY field of jaxib_grid (a plan made for the projection
// must never serve a step with another dt / nu / scheme).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define JAXIB_HAVE_XLA_FFI 1
#endif
#endif

#ifdef JAXIB_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "../../include/jaxib_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;
using U8 = ffi::Buffer<ffi::U8>;
using RF32 = ffi::ResultBuffer<ffi::F32>;

ffi::Error Status(int rc) {
  if (rc == JAXIB_OK) return ffi::Error::Success();
  const ffi::ErrorCode code = rc == JAXIB_ERR_INVALID       ? ffi::ErrorCode::kInvalidArgument
                              : rc == JAXIB_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                              : rc == JAXIB_ERR_SCRATCH     ? ffi::ErrorCode::kResourceExhausted
                                                            : ffi::ErrorCode::kInternal;
  return ffi::Error(code, jaxib_last_error());
}

template <typename T>
void Get(const ffi::Dictionary& a, const char* name, T* out) {
  if (!a.contains(name)) return;
  auto v = a.get<T>(name);
  if (v.has_value()) *out = *v;
}

// struct jaxib_grid from the call's attributes; the array shape [batch?, nx, ny] comes from a field operand
jaxib_grid MakeGrid(ffi::Span<int64_t> dims, const ffi::Dictionary& a) {
  jaxib_grid g;
  memset(&g, 0, sizeof(g));
  const size_t r = dims.size();
  g.nx = static_cast<int32_t>(dims[r - 2]);
  g.ny = static_cast<int32_t>(dims[r - 1]);
  g.batch = r == 3 ? static_cast<int32_t>(dims[0]) : 1;
  g.dx = g.dy = 1.0;
  g.nu = -1.0;
  g.density = 1.0;
  g.ib_window = 6;
  g.convect = JAXIB_CONVECT_NONE;
  Get(a, "lower_x", &g.lower_x); Get(a, "lower_y", &g.lower_y);
  Get(a, "dx", &g.dx); Get(a, "dy", &g.dy); Get(a, "dt", &g.dt); Get(a, "nu", &g.nu); Get(a, "density", &g.density);
  Get(a, "bc", &g.bc); Get(a, "convect", &g.convect); Get(a, "ib_window", &g.ib_window);
  Get(a, "u_wall_lo", &g.u_wall[0]); Get(a, "u_wall_hi", &g.u_wall[1]);
  Get(a, "v_wall_lo", &g.v_wall[0]); Get(a, "v_wall_hi", &g.v_wall[1]);
  Get(a, "u_wall_x_lo", &g.u_wall_x[0]); Get(a, "u_wall_x_hi", &g.u_wall_x[1]);
  Get(a, "v_wall_x_lo", &g.v_wall_x[0]); Get(a, "v_wall_x_hi", &g.v_wall_x[1]);
  Get(a, "force_x", &g.force[0]); Get(a, "force_y", &g.force[1]);
  return g;
}

// plans are immutable once built; cached per device on the bytes of the whole jaxib_grid
const jaxib_plan* GetPlan(const jaxib_grid& g) {
  static std::mutex mu;
  static std::map<std::string, jaxib_plan*> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::string key(reinterpret_cast<const char*>(&g), sizeof(g));
  key.append(reinterpret_cast<const char*>(&dev), sizeof(dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  jaxib_plan* plan = nullptr;
  if (jaxib_plan_create(&g, &plan) != JAXIB_OK) return nullptr;
  cache[key] = plan;
  return plan;
}

// XLA may or may not honour input_output_aliases: copy only when it declined
void CopyIfDistinct(cudaStream_t s, const F32& in, RF32& out) {
  if (out->typed_data() != in.typed_data())
    cudaMemcpyAsync(out->typed_data(), in.typed_data(), in.element_count() * sizeof(float), cudaMemcpyDeviceToDevice, s);
}

jaxib_bodies MakeBodies(const F32& x0, const F32& y0, const S32& body_offset, const F32& marker_pack, double max_radius) {
  jaxib_bodies b;
  memset(&b, 0, sizeof(b));
  b.n_markers = static_cast<int32_t>(x0.element_count());
  b.n_bodies = static_cast<int32_t>(body_offset.element_count()) - 1;
  b.body_offset = body_offset.typed_data();
  b.x0 = x0.typed_data();
  b.y0 = y0.typed_data();
  b.max_radius = max_radius;
  b.marker_pack = marker_pack.element_count() == 4 * x0.element_count() ? marker_pack.typed_data() : nullptr;
  return b;
}

// ---- equations.py:42-83 navier_stokes_explicit_terms (`convect` + `diffuse` + `forcing` plug-ins) ----------
ffi::Error ExplicitTermsImpl(cudaStream_t s, F32 u, F32 v, F32 perm, RF32 du, RF32 dv, ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(u.dimensions(), a);
  return Status(jaxib_explicit_terms(s, &g, u.typed_data(), v.typed_data(),
                                     perm.element_count() == u.element_count() ? perm.typed_data() : nullptr,
                                     du->typed_data(), dv->typed_data()));
}

// ---- time_stepping.py:191-208 predictor ---------------------------------------------------------------------
ffi::Error ExplicitPredictImpl(cudaStream_t s, F32 u, F32 v, F32 p, F32 perm, RF32 us, RF32 vs, ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(u.dimensions(), a);
  int32_t incremental = 1;
  Get(a, "incremental_pressure", &incremental);
  return Status(jaxib_explicit_predict(s, &g, u.typed_data(), v.typed_data(), incremental ? p.typed_data() : nullptr,
                                       perm.element_count() == u.element_count() ? perm.typed_data() : nullptr,
                                       us->typed_data(), vs->typed_data()));
}

// ---- convolution_functions.py:11-37 new_surf_fn (`surface_fn`) -------------------------------------------------
ffi::Error IbInterpImpl(cudaStream_t s, F32 field, F32 xp, F32 yp, RF32 out, ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(field.dimensions(), a);
  double off_x = 1.0, off_y = 0.5;
  Get(a, "offset_x", &off_x); Get(a, "offset_y", &off_y);
  return Status(jaxib_ib_interp(s, &g, off_x, off_y, field.typed_data(), xp.typed_data(), yp.typed_data(),
                                static_cast<int32_t>(xp.element_count()), out->typed_data(), nullptr, nullptr));
}

// ---- IBM_Force.py:66-87 spreading of a marker cloud; `out` aliases `base` (out = base + scale * spread) -------
ffi::Error IbSpreadImpl(cudaStream_t s, ffi::ScratchAllocator scratch, F32 base, F32 F, F32 xp, F32 yp, F32 dS, RF32 out,
                        ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(base.dimensions(), a);
  double off_x = 1.0, off_y = 0.5, scale = 1.0;
  Get(a, "offset_x", &off_x); Get(a, "offset_y", &off_y); Get(a, "scale", &scale);
  // scratch for the cell-binned gather (markers counting-sorted by owning cell)
  const size_t bytes = jaxib_ib_spread_scratch_bytes(&g, static_cast<int32_t>(xp.element_count()));
  auto mem = scratch.Allocate(bytes);
  if (!mem.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jaxib scratch");
  CopyIfDistinct(s, base, out);
  return Status(jaxib_ib_spread(s, &g, off_x, off_y, F.typed_data(), xp.typed_data(), yp.typed_data(), dS.typed_data(),
                                static_cast<int32_t>(xp.element_count()), scale, out->typed_data(), *mem, bytes));
}

// ---- IBM_Force.py:109-115 calc_IBM_force_NEW_MULTIPLE fused with time_stepping.py:223 (`IBM_forcing`) ----------
ffi::Error IbForceImpl(cudaStream_t s, ffi::ScratchAllocator scratch, F32 u_star, F32 v_star, F32 body_state, F32 x0,
                       F32 y0, S32 body_offset, F32 marker_pack, RF32 u_out, RF32 v_out, ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(u_star.dimensions(), a);
  double max_radius = 0.0;
  Get(a, "max_radius", &max_radius);
  jaxib_bodies b = MakeBodies(x0, y0, body_offset, marker_pack, max_radius);
  const size_t bytes = (size_t)5 * g.batch * b.n_markers * sizeof(float) + 256;
  auto mem = scratch.Allocate(bytes);
  if (!mem.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jaxib scratch");
  CopyIfDistinct(s, u_star, u_out);
  CopyIfDistinct(s, v_star, v_out);
  return Status(jaxib_ib_force(s, &g, &b, body_state.typed_data(), u_out->typed_data(), v_out->typed_data(), *mem, bytes));
}

// ---- pressure.py:94-154 solve_fast_diag / _moving_wall / _Far_Field (`pressure_solve`) --------------------------
ffi::Error PressureSolveImpl(cudaStream_t s, ffi::ScratchAllocator scratch, F32 u, F32 v, RF32 q, ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(u.dimensions(), a);
  const jaxib_plan* plan = GetPlan(g);
  if (!plan) return Status(JAXIB_ERR_UNSUPPORTED);
  const size_t bytes = jaxib_scratch_bytes(plan, 0);
  auto mem = scratch.Allocate(bytes);
  if (!mem.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jaxib scratch");
  return Status(jaxib_pressure_solve(s, plan, u.typed_data(), v.typed_data(), q->typed_data(), *mem, bytes));
}

// ---- pressure.py:58-91 projection_and_update_pressure ------------------------------------------------------------
ffi::Error ProjectImpl(cudaStream_t s, ffi::ScratchAllocator scratch, F32 u, F32 v, F32 p, RF32 u_out, RF32 v_out,
                       RF32 p_out, ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(u.dimensions(), a);
  const jaxib_plan* plan = GetPlan(g);
  if (!plan) return Status(JAXIB_ERR_UNSUPPORTED);
  const size_t bytes = jaxib_scratch_bytes(plan, 0);
  auto mem = scratch.Allocate(bytes);
  if (!mem.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jaxib scratch");
  return Status(jaxib_project(s, plan, u.typed_data(), v.typed_data(), p.typed_data(), u_out->typed_data(),
                              v_out->typed_data(), p_out->typed_data(), *mem, bytes));
}

// ---- time_stepping.py:144-255 step_fn (`nsteps` of them); (u, v, p) aliased in/out ------------------------------
// `perm` is the Brinkman permeability (one element: none); `walls` the host-evaluated wall values of every step,
// [nsteps, 8] float32 ON THE DEVICE is not usable (jaxib_step wants them on the host), so they come as the
// attributes u_wall_* / v_wall_* and stay constant inside one call: call once per step for moving walls.
ffi::Error StepImpl(cudaStream_t s, ffi::ScratchAllocator scratch, F32 u, F32 v, F32 p, F32 body_table, F32 x0, F32 y0,
                    S32 body_offset, F32 marker_pack, F32 perm, RF32 u_out, RF32 v_out, RF32 p_out, ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(u.dimensions(), a);
  const jaxib_plan* plan = GetPlan(g);
  if (!plan) return Status(JAXIB_ERR_UNSUPPORTED);
  double max_radius = 0.0;
  int32_t nsteps = 1, incremental = 1;
  Get(a, "max_radius", &max_radius); Get(a, "nsteps", &nsteps); Get(a, "incremental_pressure", &incremental);
  jaxib_bodies b = MakeBodies(x0, y0, body_offset, marker_pack, max_radius);
  const size_t bytes = jaxib_scratch_bytes(plan, b.n_markers);
  auto mem = scratch.Allocate(bytes);
  if (!mem.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jaxib scratch");
  CopyIfDistinct(s, u, u_out);
  CopyIfDistinct(s, v, v_out);
  CopyIfDistinct(s, p, p_out);
  return Status(jaxib_step(s, plan, b.n_markers > 0 ? &b : nullptr, body_table.typed_data(), nullptr,
                           perm.element_count() == u.element_count() ? perm.typed_data() : nullptr, nullptr, incremental,
                           nsteps, u_out->typed_data(), v_out->typed_data(), p_out->typed_data(), *mem, bytes));
}

// ---- MD/CFD_MD_coupling.py:7-27 + MD/simulate.py:81-97 tracer step; R aliased in/out ----------------------------
ffi::Error TracerStepImpl(cudaStream_t s, F32 u, F32 v, F32 R, F32 pair_force, F32 xi, U8 is_mobile, RF32 R_out,
                          ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(u.dimensions(), a);
  double dt_md = 0.0, mass_gamma = 1.0, kT = 0.0;
  Get(a, "dt_md", &dt_md); Get(a, "mass_gamma", &mass_gamma); Get(a, "kT", &kT);
  const int32_t n = static_cast<int32_t>(R.element_count() / 2);
  CopyIfDistinct(s, R, R_out);
  return Status(jaxib_tracer_step(s, &g, u.typed_data(), v.typed_data(),
                                  pair_force.element_count() == R.element_count() ? pair_force.typed_data() : nullptr,
                                  xi.element_count() == R.element_count() ? xi.typed_data() : nullptr,
                                  is_mobile.element_count() == (size_t)n ? is_mobile.typed_data() : nullptr, dt_md,
                                  mass_gamma, kT, n, R_out->typed_data()));
}

ffi::Error BilinearSampleImpl(cudaStream_t s, F32 field, F32 R, RF32 out, ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(field.dimensions(), a);
  double off_x = 1.0, off_y = 0.5;
  Get(a, "offset_x", &off_x); Get(a, "offset_y", &off_y);
  return Status(jaxib_bilinear_sample(s, &g, off_x, off_y, field.typed_data(), R.typed_data(),
                                      static_cast<int32_t>(R.element_count() / 2), out->typed_data()));
}

// ---- MD/interaction_potential.py:6-22 through smap.pair + quantity.force --------------------------------------------
ffi::Error PairForceImpl(cudaStream_t s, F32 R, S32 species, F32 h, F32 r0, RF32 F, ffi::Dictionary a) {
  double D0 = 0.0, alpha = 10.0, k = 1.0, box_x = 1.0, box_y = 1.0;
  Get(a, "D0", &D0); Get(a, "alpha", &alpha); Get(a, "k", &k); Get(a, "box_x", &box_x); Get(a, "box_y", &box_y);
  int32_t n_species = 1;
  while ((size_t)(n_species + 1) * (n_species + 1) <= h.element_count()) ++n_species;
  return Status(jaxib_pair_force(s, R.typed_data(), species.typed_data(), static_cast<int32_t>(R.element_count() / 2),
                                 n_species, h.typed_data(), r0.typed_data(), D0, alpha, k, box_x, box_y, F->typed_data()));
}

// ---- penalty/util_funs.py:26-123 permeability mask ----------------------------------------------------------------
ffi::Error PenaltyPermImpl(cudaStream_t s, F32 like, F32 centers, F32 Rtheta, RF32 perm, ffi::Dictionary a) {
  jaxib_grid g = MakeGrid(like.dimensions(), a);
  double K = 0.0, width = 1.0;
  Get(a, "K", &K); Get(a, "width", &width);
  const int32_t nb = static_cast<int32_t>(centers.element_count() / 2);
  return Status(jaxib_penalty_perm(s, &g, centers.typed_data(), Rtheta.typed_data(), nb,
                                   nb > 0 ? static_cast<int32_t>(Rtheta.element_count() / nb) : 0, K, width,
                                   perm->typed_data()));
}

}  // namespace

#define JAXIB_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibExplicitTerms, ExplicitTermsImpl,
                              JAXIB_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibExplicitPredict, ExplicitPredictImpl,
                              JAXIB_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibIbInterp, IbInterpImpl, JAXIB_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibIbSpread, IbSpreadImpl,
                              JAXIB_STREAM.Ctx<ffi::ScratchAllocator>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibIbForce, IbForceImpl,
                              JAXIB_STREAM.Ctx<ffi::ScratchAllocator>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<S32>().Arg<F32>().Ret<F32>().Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibPressureSolve, PressureSolveImpl,
                              JAXIB_STREAM.Ctx<ffi::ScratchAllocator>().Arg<F32>().Arg<F32>().Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibProject, ProjectImpl,
                              JAXIB_STREAM.Ctx<ffi::ScratchAllocator>().Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>()
                                  .Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibStep, StepImpl,
                              JAXIB_STREAM.Ctx<ffi::ScratchAllocator>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<S32>().Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibTracerStep, TracerStepImpl,
                              JAXIB_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<U8>().Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibBilinearSample, BilinearSampleImpl, JAXIB_STREAM.Arg<F32>().Arg<F32>().Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibPairForce, PairForceImpl,
                              JAXIB_STREAM.Arg<F32>().Arg<S32>().Arg<F32>().Arg<F32>().Ret<F32>().Attrs());
XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibPenaltyPerm, PenaltyPermImpl, JAXIB_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>().Attrs());
// Not wrapped on purpose: jaxib_plan_* (internal to the handlers above), the slab-decomposition stage entry points
// and jaxib_slab_step (driven by the library's own multi-process host side, jax_ib_b200/slab.py -- XLA does not hand
// peer mappings or communicators to FFI handlers) and jaxib_step_profile (a measurement aid that synchronises).
#endif  // JAXIB_HAVE_XLA_FFI
