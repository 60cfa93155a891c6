// Kernel launch helper: every kernel of the library is launched with the programmatic-dependent-launch attribute, so
// kernel N+1 can be scheduled (and run its prologue / weight prefetch) while kernel N drains. Kernels call
// grid_dep_launch() early and grid_dep_wait() before touching memory produced by their predecessor.
// OPUS_PDL=0 in the environment (read when the context is created) falls back to plain stream order, OPUS_PDL=2 enables
// it everywhere (A/B measurements).
#pragma once
#include <cuda_runtime.h>
#include <utility>

#include "context.h"

namespace opus {

// `small` = decode-sized launch. Measured on B200: PDL shortens the (HBM-bound, launch-latency-sensitive) decode step
// but slows the power-capped multi-wave prefill/encoder kernels by a few percent, so only small launches opt in
// (OPUS_PDL=2 at context creation forces it for every launch, OPUS_PDL=0 disables it).
inline bool pdl_for(bool small) {
  const int mode = ctx().tun.pdl;
  return mode == 2 || (mode == 1 && small);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool small, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_for(small) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace opus
