// Kernel launch helper: every kernel of the library is launched with the programmatic-dependent-launch attribute, so
// kernel N+1 can be scheduled (and run its prologue / weight prefetch) while kernel N drains. Kernels call
// grid_dep_launch() early and grid_dep_wait() before touching memory produced by their predecessor.
// OPUS_PDL=0 in the environment falls back to plain stream order (for A/B measurements).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>

namespace opus {

inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = std::getenv("OPUS_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace opus
