// GPU probe (not part of the library): issue cost of tcgen05.mma (kind::f16, M = 128, K = 16, SS operands) as a function of
// N, operand major-ness and swizzle width, with one and two CTAs per SM -- the cost model DESIGN.md section 4 quotes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<pkg>/csrc -o tools/mma_probe tools/mma_probe.cu
//   ./tools/mma_probe
// One warp per CTA issues `reps` accumulating MMAs back to back on zeroed shared memory (elected lane, as the kernels do),
// commits to an mbarrier and waits; cycles = clock64 around issue + completion.  Operand values do not matter.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

__global__ void __launch_bounds__(64, 2) probe(int N, int a_mn, int b_mn, int wa, int wb, int reps, long long* out,
                                               int a_shift_rows, int fill) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 80 * 1024 / 16; i += blockDim.x) {
    // fill = 1: small non-zero halves (0x2c00 | bits = values around 0.06) instead of zeros
    const uint32_t v = fill ? (0x2c002c00u | ((uint32_t)(i * 2654435761u) & 0x03ff03ffu)) : 0u;
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(v, v ^ 0x00110011u, v ^ 0x01010101u, v);
  }
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  uint32_t cols = 32;
  while (cols < (uint32_t)N) cols <<= 1;
  if (warp == 0) tc::tmem_alloc(&tmem_ptr, cols);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1) {
    auto lay = [](int w) { return w == 128 ? tc::SWZ_128B : (w == 64 ? tc::SWZ_64B : tc::SWZ_32B); };
    // a_shift_rows: the A start address moved by whole rows (the halo kernels' dw / dh tap shifts: not atom-aligned)
    const uint32_t a_base = tc::smem_u32(smem) + (uint32_t)(a_shift_rows * wa), b_base = tc::smem_u32(smem) + 40 * 1024;
    // K-major: rows of `w` bytes, 8-row groups w*8 apart (SBO); MN-major: chunks of w bytes of M/N, 128 K-rows... as in
    // conv_tc3.cu (K-major A / B) and wgrad_tc.cu (MN-major A / B): LBO = stride between w-byte chunks, SBO = 8 rows
    const uint64_t ad = a_mn ? tc::smem_desc(a_base, 16 * wa, 8 * wa, lay(wa)) : tc::smem_desc(a_base, 16, 8 * wa, lay(wa));
    const uint64_t bd = b_mn ? tc::smem_desc(b_base, 16 * wb, 8 * wb, lay(wb)) : tc::smem_desc(b_base, 16, 8 * wb, lay(wb));
    const uint32_t idesc = tc::idesc_f16(128, N, a_mn, b_mn);
    // K advance inside a swizzle row: +32 bytes per K = 16 step (K-major); MN-major: +16 rows
    const uint32_t ka = a_mn ? (16 * wa) >> 4 : 2, kb = b_mn ? (16 * wb) >> 4 : 2;
    const int ksteps_a = a_mn ? 4 : wa / 32, ksteps_b = b_mn ? 4 : wb / 32;
    __syncwarp();
    const long long t0 = clock64();
    // descriptors of the (up to) four K steps precomputed: the loop body is the MMA issue alone, as in the kernels
    uint64_t a4[4], b4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a4[k] = ad + (uint64_t)((k & (ksteps_a - 1)) * ka);
      b4[k] = bd + (uint64_t)((k & (ksteps_b - 1)) * kb);
    }
    for (int r = 0; r < reps; r += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) tc::mma_f16_ss_elect(tmem, a4[k], b4[k], idesc, 1u);
    }
    const long long t1 = clock64();
    tc::mma_commit_elect(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, cols);
}

// Tile structure of the conv kernels: batches of `batch` MMAs, each batch committed to an mbarrier; before issuing batch i
// the issuing warp waits for the completion of batch i - 2 (a double-buffered accumulator handed back by an epilogue
// that takes no time) and runs the tcgen05 fence the kernels run.  Cycles per MMA over `nb` batches.
__global__ void __launch_bounds__(64, 2) probe_batched(int N, int batch, int nb, int wait_prev, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 80 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); tc::fence_barrier_init(); }
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * N)) cols <<= 1;
  if (warp == 0) tc::tmem_alloc(&tmem_ptr, cols);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  if (warp == 1) {
    const uint32_t a_base = tc::smem_u32(smem), b_base = a_base + 40 * 1024;
    const uint64_t ad = tc::smem_desc(a_base, 16, 8 * 64, tc::SWZ_64B), bd = tc::smem_desc(b_base, 16, 8 * 64, tc::SWZ_64B);
    const uint32_t idesc = tc::idesc_f16(128, N, 0, 0);
    uint32_t phase[2] = {0, 0};
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < nb; ++i) {
      const int s = i & 1;
      if (i >= 2 && wait_prev) {                       // accumulator s was last written by batch i - 2
        tc::mbar_wait(&bar[s], phase[s]);
        phase[s] ^= 1;
        tc::fence_after_sync();
      }
      for (int r = 0; r < batch; r += 2) {
        tc::mma_f16_ss_elect(tmem + s * N, ad + (uint64_t)((r % 18) * 4), bd, idesc, r != 0);
        tc::mma_f16_ss_elect(tmem + s * N, ad + (uint64_t)((r % 18) * 4 + 2), bd + 2, idesc, 1u);
      }
      tc::mma_commit_elect(&bar[s]);
    }
    const long long t1 = clock64();
    if (!wait_prev) {                                  // drain: barriers completed several phases, just wait a while
      for (int k = 0; k < 2000; ++k) asm volatile("nanosleep.u32 20;");
    } else {
      for (int s = 0; s < 2; ++s) { tc::mbar_wait(&bar[(nb + s) & 1], phase[(nb + s) & 1]); }
    }
    const long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, cols);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  struct Cfg { const char* name; int a_mn, b_mn, wa, wb; };
  const Cfg cfgs[] = {
      {"A K-major SW128, B K-major SW128 (conv fwd, K >= 64)", 0, 0, 128, 128},
      {"A K-major SW64 , B K-major SW64  (conv fwd, K = 32)", 0, 0, 64, 64},
      {"A K-major SW32 , B K-major SW32  (conv fwd, K = 16)", 0, 0, 32, 32},
      {"A MN-major SW128, B MN-major SW128 (wgrad, 64-ch slabs)", 1, 1, 128, 128},
      {"A MN-major SW64 , B MN-major SW64  (wgrad, 32-ch slabs)", 1, 1, 64, 64},
      {"A MN-major SW64 , B MN-major SW128", 1, 1, 64, 128},
      {"A MN-major SW128, B MN-major SW64", 1, 1, 128, 64},
  };
  const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
  const int reps = 512;
  for (int var = 0; var < 8; ++var) {
    const int two = var & 1, shift = (var >> 1) == 1 ? 1 : ((var >> 1) == 2 ? 11 : 0), fill = (var >> 1) == 3;
    printf("---- %s (grid %d), A start shifted by %d rows, %s data, %d MMAs per CTA: cycles per MMA (issue only / issue + completion)\n",
           two ? "two CTAs per SM" : "one CTA per SM", two ? 296 : 148, shift, fill ? "non-zero" : "zero", reps);
    for (const Cfg& c : cfgs) {
      if (shift && c.a_mn) continue;
      printf("%-58s", c.name);
      for (int N : Ns) {
        long long h[2] = {0, 0};
        for (int it = 0; it < 2; ++it) {
          probe<<<two ? 296 : 148, 64, 90 * 1024>>>(N, c.a_mn, c.b_mn, c.wa, c.wb, reps, d, shift, fill);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf(" [N=%d: %s]", N, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("  N=%-3d %5.1f/%5.1f", N, (double)h[0] / reps, (double)h[1] / reps);
      }
      printf("\n");
    }
  }
  cudaFuncSetAttribute(probe_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int two = 0; two < 2; ++two) {
    for (int wait_prev = 0; wait_prev < 2; ++wait_prev) {
      printf("---- batched issue, %s, %s: cycles per MMA (issue loop)\n", two ? "two CTAs per SM" : "one CTA per SM",
             wait_prev ? "wait for batch i-2 + fence before batch i" : "commit only");
      for (int N : {32, 64, 96, 128, 256}) {
        printf("N=%-3d", N);
        for (int batch : {4, 6, 18, 36, 72, 144}) {
          const int nb = 1152 / batch / (N >= 128 ? 2 : 1);
          long long h[2] = {0, 0};
          for (int it = 0; it < 2; ++it) {
            probe_batched<<<two ? 296 : 148, 64, 90 * 1024>>>(N, batch, nb, wait_prev, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf(" [%s]", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("   batch %3d: %5.1f", batch, (double)h[0] / (nb * batch));
        }
        printf("\n");
      }
    }
  }
  return 0;
}
