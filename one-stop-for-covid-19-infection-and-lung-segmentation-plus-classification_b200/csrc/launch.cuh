// Kernel-launch helper: counts launches (bench.py's "gpu_launches") and converts launch errors into the C-ABI's int
// return.  Optionally (b2u_set_option("pdl", 1)) every kernel is launched with programmatic dependent launch: the
// next kernel of the stream / graph may be scheduled while the current one drains (its CTAs run B2U_PDL_PROLOGUE
// first: `griddepcontrol.wait` blocks until the preceding grid has completed and its writes are visible, so data
// dependencies are unchanged).  Measured on the U-Net 512x512 step (B200, CUDA graph): 5.49 ms with PDL against
// 5.26 ms without -- early-resident dependent CTAs cost the persistent kernels more than the overlapped launch
// latency gains -- so it is OFF by default; without the launch attribute the two instructions are no-ops.
#pragma once
#include "common.cuh"

extern unsigned long long g_b2u_launches;
extern int g_b2u_pdl;          // b2u_set_option("pdl", 0/1)

// first statements of every kernel: let the dependent grid start early, then wait for the grid we depend on
#define B2U_PDL_LAUNCH_DEPENDENTS() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define B2U_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define B2U_PDL_PROLOGUE()        \
  do {                            \
    B2U_PDL_LAUNCH_DEPENDENTS();  \
    B2U_PDL_WAIT();               \
  } while (0)

template <typename... KArgs, typename... Args>
static inline cudaError_t b2u_launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                        Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_b2u_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// the same with thread-block clusters of `cluster_x` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
static inline cudaError_t b2u_launch_cluster_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                                cudaStream_t stream, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster_x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_b2u_pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#define B2U_LAUNCH_CLUSTER(kern, grid, block, smem, stream, cluster_x, ...)                          \
  do {                                                                                               \
    auto _kfn = kern;                                                                                \
    cudaError_t _le = b2u_launch_cluster_ex(_kfn, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream), \
                                            (cluster_x), __VA_ARGS__);                               \
    __atomic_fetch_add(&g_b2u_launches, 1ULL, __ATOMIC_RELAXED);                                     \
    B2U_CHECK_CUDA(_le);                                                                             \
    B2U_LAUNCH_CHECK();                                                                              \
  } while (0)

#define B2U_LAUNCH(kern, grid, block, smem, stream, ...)                                             \
  do {                                                                                               \
    auto _kfn = kern;                                                                                \
    cudaError_t _le = b2u_launch_ex(_kfn, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream), \
                                    __VA_ARGS__);                                                    \
    __atomic_fetch_add(&g_b2u_launches, 1ULL, __ATOMIC_RELAXED);                                     \
    B2U_CHECK_CUDA(_le);                                                                             \
    B2U_LAUNCH_CHECK();                                                                              \
  } while (0)
