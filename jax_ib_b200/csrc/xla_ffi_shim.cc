// XLA-FFI bindings of the C ABI (include/jaxib_b200.h) -- the binding a jax_ib maintainer adds so that
// the reference's host side stays Python/JAX.
//
// jaxlib's headers (xla/ffi/api/ffi.h) are NOT present in this image, so this file is
// compile-guarded: it builds to an empty object here and is not part of libjaxib_b200.so.  It is pure
// argument plumbing over the tested C entry points: XLA owns the buffers (inputs immutable, outputs
// pre-allocated, in-place only through input_output_aliases) and hands us its CUDA stream and a
// scratch allocator; handlers may be invoked concurrently from several host threads (one per
// device), which is safe because the library keeps no mutable globals besides the plan cache below.
//
//   build (where jaxlib is installed):
//     nvcc -std=c++17 -shared -Xcompiler -fPIC -I$(python -c "import jaxlib, os; print(os.path.join(os.path.dirname(jaxlib.__file__), 'include'))") \
//          xla_ffi_shim.cc -L../_native -ljaxib_b200 -o libjaxib_b200_ffi.so
//   register (Python):
//     jax.ffi.register_ffi_target("jaxib_step", jax.ffi.pycapsule(lib.JaxibStep), platform="CUDA")
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define JAXIB_HAVE_XLA_FFI 1
#endif
#endif

#ifdef JAXIB_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <tuple>

#include "../../include/jaxib_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Status(int rc) {
  if (rc == JAXIB_OK) return ffi::Error::Success();
  return ffi::Error(rc == JAXIB_ERR_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    jaxib_last_error());
}

jaxib_grid MakeGrid(ffi::Span<const int64_t> dims, double lx, double ly, double dx, double dy, double dt, double nu,
                    double density, int32_t bc, int32_t convect, int32_t window) {
  jaxib_grid g{};
  const size_t r = dims.size();
  g.nx = static_cast<int32_t>(dims[r - 2]);
  g.ny = static_cast<int32_t>(dims[r - 1]);
  g.batch = r == 3 ? static_cast<int32_t>(dims[0]) : 1;
  g.lower_x = lx; g.lower_y = ly; g.dx = dx; g.dy = dy; g.dt = dt; g.nu = nu; g.density = density;
  g.bc = bc; g.convect = convect; g.ib_window = window;
  return g;
}

// plans are immutable once built; cached per (device, grid) behind a mutex
const jaxib_plan* GetPlan(const jaxib_grid& g) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int, int, int, double, double>, jaxib_plan*> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  auto key = std::make_tuple(dev, g.nx, g.ny, g.batch, g.bc, g.dx, g.dy);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  jaxib_plan* plan = nullptr;
  if (jaxib_plan_create(&g, &plan) != JAXIB_OK) return nullptr;
  cache[key] = plan;
  return plan;
}

// pressure.py:58-91 projection_and_update_pressure as one custom call
ffi::Error ProjectImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> u,
                       ffi::Buffer<ffi::F32> v, ffi::Buffer<ffi::F32> p, ffi::ResultBuffer<ffi::F32> u_out,
                       ffi::ResultBuffer<ffi::F32> v_out, ffi::ResultBuffer<ffi::F32> p_out, double dx, double dy,
                       int32_t bc) {
  jaxib_grid g = MakeGrid(u.dimensions(), 0, 0, dx, dy, 0, -1, 1, bc, JAXIB_CONVECT_NONE, 6);
  const jaxib_plan* plan = GetPlan(g);
  if (!plan) return Status(JAXIB_ERR_UNSUPPORTED);
  const size_t bytes = jaxib_scratch_bytes(plan, 0);
  auto mem = scratch.Allocate(bytes);
  if (!mem.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jaxib scratch");
  return Status(jaxib_project(stream, plan, u.typed_data(), v.typed_data(), p.typed_data(), u_out->typed_data(),
                              v_out->typed_data(), p_out->typed_data(), *mem, bytes));
}

// time_stepping.py:144-255 step_fn (`nsteps` of them) as one custom call; (u, v, p) are aliased in/out
ffi::Error StepImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> u,
                    ffi::Buffer<ffi::F32> v, ffi::Buffer<ffi::F32> p, ffi::Buffer<ffi::F32> body_table,
                    ffi::Buffer<ffi::F32> x0, ffi::Buffer<ffi::F32> y0, ffi::Buffer<ffi::S32> body_offset,
                    ffi::ResultBuffer<ffi::F32> u_out, ffi::ResultBuffer<ffi::F32> v_out,
                    ffi::ResultBuffer<ffi::F32> p_out, double lx, double ly, double dx, double dy, double dt, double nu,
                    double density, double max_radius, int32_t bc, int32_t convect, int32_t window, int32_t nsteps) {
  jaxib_grid g = MakeGrid(u.dimensions(), lx, ly, dx, dy, dt, nu, density, bc, convect, window);
  const jaxib_plan* plan = GetPlan(g);
  if (!plan) return Status(JAXIB_ERR_UNSUPPORTED);
  jaxib_bodies b{};
  b.n_markers = static_cast<int32_t>(x0.element_count());
  b.n_bodies = static_cast<int32_t>(body_offset.element_count()) - 1;
  b.body_offset = body_offset.typed_data(); b.x0 = x0.typed_data(); b.y0 = y0.typed_data(); b.max_radius = max_radius;
  const size_t bytes = jaxib_scratch_bytes(plan, b.n_markers);
  auto mem = scratch.Allocate(bytes);
  if (!mem.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jaxib scratch");
  // outputs alias inputs (input_output_aliases={0:0, 1:1, 2:2}); copy only if XLA declined the alias
  const size_t n = u.element_count() * sizeof(float);
  if (u_out->typed_data() != u.typed_data()) cudaMemcpyAsync(u_out->typed_data(), u.typed_data(), n, cudaMemcpyDeviceToDevice, stream);
  if (v_out->typed_data() != v.typed_data()) cudaMemcpyAsync(v_out->typed_data(), v.typed_data(), n, cudaMemcpyDeviceToDevice, stream);
  if (p_out->typed_data() != p.typed_data()) cudaMemcpyAsync(p_out->typed_data(), p.typed_data(), n, cudaMemcpyDeviceToDevice, stream);
  return Status(jaxib_step(stream, plan, b.n_markers ? &b : nullptr, body_table.typed_data(), nullptr, 1, nsteps,
                           u_out->typed_data(), v_out->typed_data(), p_out->typed_data(), *mem, bytes));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibProject, ProjectImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<double>("dx").Attr<double>("dy").Attr<int32_t>("bc"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(JaxibStep, StepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<double>("lower_x").Attr<double>("lower_y").Attr<double>("dx").Attr<double>("dy")
                                  .Attr<double>("dt").Attr<double>("nu").Attr<double>("density").Attr<double>("max_radius")
                                  .Attr<int32_t>("bc").Attr<int32_t>("convect").Attr<int32_t>("ib_window")
                                  .Attr<int32_t>("nsteps"));
#endif  // JAXIB_HAVE_XLA_FFI
