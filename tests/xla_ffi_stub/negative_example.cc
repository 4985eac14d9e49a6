#include <cuda_runtime.h>
#include "xla/ffi/api/ffi.h"
namespace ffi = xla::ffi;
ffi::Error Bad(cudaStream_t s, ffi::Buffer<ffi::F32> a, double wrong) { return ffi::Error::Success(); }
XLA_FFI_DEFINE_HANDLER_SYMBOL(BadSym, Bad, ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<ffi::F32>>().Attrs());
