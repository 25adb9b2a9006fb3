// Kernel-launch helper: counts launches (bench.py's "gpu_launches") and converts launch errors into
// the C-ABI's int return.
#pragma once
#include "common.cuh"

extern unsigned long long g_b2u_launches;

#define B2U_LAUNCH(kern, grid, block, smem, stream, ...)                                   \
  do {                                                                                     \
    auto _kfn = kern;                                                                      \
    _kfn<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                \
    __atomic_fetch_add(&g_b2u_launches, 1ULL, __ATOMIC_RELAXED);                           \
    B2U_LAUNCH_CHECK();                                                                    \
  } while (0)
