// Programmatic dependent launch (PDL) between the ~100 kernels of a timestep.
//
// Every kernel of the per-timestep sequence is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may be
// scheduled while the previous kernel is still draining, run their set-up (mbarrier init, TMEM allocation, tensor-map prefetch,
// loads of CONSTANT data such as weights) and then block in `griddepcontrol.wait` until the previous grid has completed and its
// memory is visible.  Persistent kernels (all CTAs resident) release their dependents early with `griddepcontrol.launch_dependents`;
// for multi-wave kernels the release is implicit when the last block exits.  Inside stream capture the attribute becomes a
// programmatic edge of the CUDA graph.  `griddepcontrol.wait` is a no-op for a kernel launched without the attribute, so every
// kernel carries it unconditionally.  MEASURED SLOWER on the bench workload (see pdl_flag() in ld_engine.cu), hence OFF by default:
// `ld_set_option(h, "pdl", 1)` or env LD_PDL=1 enables it.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

namespace ld {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// process-wide switch (ld_set_option "pdl"; initial value from env LD_PDL, default off); defined in ld_engine.cu
int& pdl_flag();
inline bool pdl_enabled() { return pdl_flag() != 0; }
// Selective mode (flag == 2): only the launch that FOLLOWS a tiny kernel (gn_coef, la_fold: a handful of blocks, 3 .. 17 us) carries the
// attribute -- its CTAs find the SMs empty (the big kernel before the tiny one has completed), start together and overlap their set-up with
// the tiny kernel, without the skew that costs the all-kernels mode its gain.  The tiny kernels raise the hint, the next launch_k consumes it.
inline int& pdl_after_small() { static thread_local int v = 0; return v; }

// kernel<<<grid, block, smem, s>>>(args...) with the PDL attribute (pdl == true and LD_PDL != 0)
// pdl: 0 never; 1 (true) eligible.  (Letting the tiny kernels themselves start early -- and release their dependents only after their own
// wait -- was measured: it cancels the gain of the selective mode, 4.666 vs 4.662 ms per timestep.)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  const int mode = pdl_flag();
  cfg.numAttrs = (pdl && (mode == 1 || (mode == 2 && pdl_after_small()))) ? 1 : 0;
  pdl_after_small() = 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace ld
