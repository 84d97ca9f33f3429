// Kernel launch helper: every kernel of the library is launched with programmatic dependent launch (PDL) enabled, so
// that its launch latency and prologue (barrier init, TMEM allocation, tensor-map prefetch, index math) overlap the
// tail of the kernel before it on the stream.  A training micro-step is ~1200 dependent launches; inside a CUDA graph
// the attribute becomes a programmatic edge.  Contract for every __global__ function launched through here:
//   * call pdl_wait() before the first access to global memory another kernel may have written (or may still read),
//   * call pdl_launch_dependents() as early as it likes (the dependent still waits for this grid's completion and
//     memory flush in its own pdl_wait()).
// ONIRIS_PDL=0 in the environment turns the attribute off for the whole process (plain stream order) for A/B
// measurements; ob_set_pdl() switches it at run time and returns the previous setting.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <utility>

namespace ob {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Process-wide switch (ob_set_pdl).  The backward pass turns it off: there the weight-gradient kernels run on a second
// stream, and dependents that become resident early (holding shared memory / TMEM while they wait) take SMs the other
// stream's CTAs would have used -- measured: forward 6.05 -> 5.48 ms with PDL, two-stream backward 11.0 -> 11.4 ms.
// Mode 2 (only the light, elementwise kernels) was measured too and is no better there (11.29 ms): the backward pass
// runs with PDL off.
inline std::atomic<int>& pdl_flag() {
  static std::atomic<int> flag{-1};
  return flag;
}
// 0: off, 1: every kernel, 2: only "light" kernels (no big shared-memory / TMEM footprint: the elementwise ones)
inline int pdl_mode() {
  int v = pdl_flag().load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("ONIRIS_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
    pdl_flag().store(v, std::memory_order_relaxed);
  }
  return v;
}

// cluster_x / cluster_y > 1 launches clusters; a kernel with more than 48 KB of dynamic shared memory counts as "heavy".
template <typename... KArgs, typename... Args>
inline cudaError_t launch_xy(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                             int cluster_y, Args&&... args) {
  const int mode = pdl_mode();
  const bool pdl = mode == 1 || (mode == 2 && smem <= 48 * 1024);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1 || cluster_y > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = cluster_y; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                          Args&&... args) {
  return launch_xy(kern, grid, block, smem, st, cluster_x, 1, std::forward<Args>(args)...);
}

}  // namespace ob
