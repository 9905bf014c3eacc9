// rsrx_pdl.cuh — programmatic dependent launch for the trainers' kernel chains.
//
// A PPO minibatch step is ~35 short launches (5-25 us each) replayed from a CUDA graph; between two dependent kernels the GPU
// otherwise drains the first grid completely before it starts scheduling the second.  Launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, a kernel may be scheduled as soon as every CTA of its predecessor has
// executed griddepcontrol.launch_dependents (first statement of every kernel here); it then runs its own prologue (barrier
// initialisation, TMEM allocation, tensor-map prefetch, shared-memory zeroing) and blocks in griddepcontrol.wait until the
// predecessor has COMPLETED and its writes are visible.  Rules kept by every kernel launched through launch_pdl():
//   * wait() is executed unconditionally by every thread before the first global-memory access (reads AND writes: a
//     successor may overwrite what the predecessor still reads);
//   * nothing before wait() depends on global memory.
// Both instructions are no-ops in a kernel launched without the attribute (RSRX_PDL=0, or any plain <<<>>> launch).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>

namespace rsrx {
namespace pdl {

__device__ __forceinline__ void launch_dependents() {
#ifdef RSRX_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool enabled() {
  static const bool on = [] { const char* e = getenv("RSRX_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace pdl
}  // namespace rsrx
