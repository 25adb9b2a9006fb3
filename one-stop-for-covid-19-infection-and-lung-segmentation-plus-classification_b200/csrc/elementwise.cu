// HBM-bound kernels of the U-Net step: BatchNorm statistics/apply/backward, 2x2 max-pool (+dropout),
// dropout, concat slice copies, the 1x1+sigmoid head with the BCE+Dice loss, Adam, batch gather.
// All are coalesced 8-channel-vector streaming kernels with grids sized in multiples of the SM count.
#include "common.cuh"
#include "internal.h"
#include "launch.cuh"

namespace {

constexpr int kThreads = 256;

// Thread mapping shared by the streaming kernels: a thread owns one 8-channel group `g` (fixed for its
// lifetime, so per-channel constants live in registers) and walks pixels `p` with a grid stride --
// no integer division in the loop, 16-byte (fp16) / 32-byte (fp32) accesses, consecutive threads on
// consecutive addresses.
#define PIXEL_LANE_LOOP(C, npix)                                                          \
  const int cg = (C) >> 3;                                                                \
  const int lanes = kThreads / cg;                                                        \
  const int g = threadIdx.x % cg, lane_ = threadIdx.x / cg;                               \
  if (lane_ < lanes)                                                                      \
    for (long long p = (long long)blockIdx.x * lanes + lane_; p < (npix); p += (long long)gridDim.x * lanes)

__host__ int lane_grid(long long npix, int c, int waves = 8) {
  int lanes = kThreads / (c >> 3);
  long long b = (npix + lanes - 1) / lanes;
  long long cap = (long long)B2U_NUM_SMS * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__host__ int stream_grid(long long work_items, int per_block = kThreads, int waves = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)B2U_NUM_SMS * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ __forceinline__ void keep_factors(float p, uint64_t e0, const b2u_step_state* st, int op_id, float f[8]);
// 8 channels kept PACKED (as loaded) in registers and unpacked on demand
template <typename T> struct Pack8;
template <> struct Pack8<__half> {
  uint4 u;
  __device__ __forceinline__ void load(const __half* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float v[8]) const {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};
template <> struct Pack8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void get(float v[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// dropout factors of 8 consecutive elements from their stored keep bits
__device__ __forceinline__ void bits_factors(unsigned b, float sc, float f[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = ((b >> i) & 1u) ? sc : 0.f;
}

// ------------------------------------------------------------------------------------------
// BatchNorm statistics: sums[c] += sum x, sums[C + c] += sum x^2   (double accumulators)
// kDrop: the BN input is dropout(x) of a Dropout layer that is NOT materialised (Conv2D -> Dropout -> BatchNormalization,
// UPP:874-876 ...): the keep mask is regenerated from the Philox stream (element index pix * C + c of the dropout
// layer's tensor, as dropout_kernel does) and x * keep / (1 - p) takes the place of x.
// ------------------------------------------------------------------------------------------
template <typename T, bool kBwd, bool kDrop>
__global__ void __launch_bounds__(kThreads, kDrop ? 2 : 4) bn_reduce_kernel(const T* __restrict__ a, int lda,
                                                             const T* __restrict__ x, int ldx, int C,
                                                             long long npix, const float* __restrict__ mean,
                                                             const float* __restrict__ invstd,
                                                             double* __restrict__ sums, int sq_off,
                                                             const b2u_step_state* __restrict__ st, float p_drop,
                                                             int op_id, uint8_t* __restrict__ drop_bits) {
  // drop_bits (bit pix * C + c, one byte per thread and pixel): the forward statistics pass runs Philox ONCE and stores the
  // keep mask; every later pass over the same tensor (apply, backward reduce / apply) reads one byte per 8 elements
  // instead of regenerating it (regenerating in all four passes cost what the removed Dropout passes had cost)
  B2U_PDL_PROLOGUE();
  // kBwd == false: a = x (stats of a).  kBwd == true: a = dy, x = bn input; sums of dy and dy*xhat.
  // block-level partial sums in fp32 (native shared-memory atomics; fp64 shared atomics are CAS loops and
  // cost more than the streaming loop), one fp64 global atomic per channel per block
  extern __shared__ float sacc[];   // [2*C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cg = C >> 3;
  const int lanes = kThreads / cg;              // pixel lanes per block
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  {
    // kBwd accumulates sum dy*x and applies  sum dy*xhat = invstd * (sum dy*x - mean * sum dy)  once per thread
    // (a thread's partial sums are short, so the subtraction loses nothing): no per-channel constants live in
    // the loop, which keeps the kernel at 4 blocks per SM
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    if (lane < lanes) {
    // two pixel groups per trip, both inside one contiguous span of the block (coalesced, 2-4 loads in flight)
    const long long stride = (long long)gridDim.x * lanes * 2;
    for (long long p = (long long)blockIdx.x * lanes * 2 + lane; p < npix; p += stride) {
      const long long p2 = p + lanes;                       // the block covers 2*lanes contiguous pixels per trip
      const bool two = p2 < npix;
      float v[2][8], xv[2][8];
      load8<T>(a + p * lda + g * 8, v[0]);
      if (two) load8<T>(a + p2 * lda + g * 8, v[1]);
      if (kBwd) {
        load8<T>(x + p * ldx + g * 8, xv[0]);
        if (two) load8<T>(x + p2 * ldx + g * 8, xv[1]);
      }
      if (kDrop) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u == 1 && !two) break;
          float f[8];
          const uint64_t e0 = (uint64_t)(u ? p2 : p) * C + g * 8;
          if (kBwd) {
            bits_factors(drop_bits[e0 >> 3], 1.f / (1.f - p_drop), f);
          } else {
            keep_factors(p_drop, e0, st, op_id, f);
            unsigned b = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) b |= (f[i] != 0.f ? 1u : 0u) << i;
            drop_bits[e0 >> 3] = (uint8_t)b;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { if (kBwd) xv[u][i] *= f[i]; else v[u][i] *= f[i]; }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        if (kBwd) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { s1[i] += v[u][i]; s2[i] = fmaf(v[u][i], xv[u][i], s2[i]); }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) { s1[i] += v[u][i]; s2[i] += v[u][i] * v[u][i]; }
        }
      }
    }
    if (kBwd) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s2[i] = invstd[g * 8 + i] * (s2[i] - mean[g * 8 + i] * s1[i]);
    }
    }
    if (lane < lanes || (cg & (cg - 1)) == 0) {      // (all threads are pixel lanes when cg is a power of two)
      group_add8(sacc, s1, cg, g);
      group_add8(sacc + C, s2, cg, g);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&sums[i], (double)sacc[i]);
    atomicAdd(&sums[sq_off + i], (double)sacc[C + i]);
  }
}

// sums[c] = colsum[c];  sums[C + c] = (sum_{tap,co} W[tap][c][co] * dW[tap][c][co] - beta[c] * colsum[c]) / gamma[c]
// one block per input channel c (see include/b200unet.h: the adjoint identity of the convolution)
__global__ void __launch_bounds__(128) bn_bwd_sums_wgrad_kernel(const float* __restrict__ w, const float* __restrict__ dw,
                                                                const float* __restrict__ colsum,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, double* __restrict__ sums,
                                                                int C, int cout, int taps) {
  B2U_PDL_PROLOGUE();
  __shared__ double part[4];
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < taps * cout; i += blockDim.x) {
    const int t = i / cout, co = i - t * cout;
    const long long idx = ((long long)t * C + c) * cout + co;
    acc += (double)w[idx] * (double)dw[idx];
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double tot = part[0] + part[1] + part[2] + part[3];
    const double s1 = (double)colsum[c];
    const double g = (double)gamma[c];
    sums[c] += s1;
    sums[C + c] += fabs(g) > 1e-12 ? (tot - (double)beta[c] * s1) / g : 0.0;
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ mmean, float* __restrict__ mvar, float momentum,
                                   float eps, int training, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd, int C) {
  B2U_PDL_PROLOGUE();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    double m = sums[c] / (double)count;
    double v = sums[C + c] / (double)count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    // Keras 2.3 normalization.py: sample variance n/(n-(1+eps)) feeds the moving average
    double unb = v * ((double)count / ((double)count - (1.0 + (double)eps)));
    mmean[c] = momentum * mmean[c] + (1.f - momentum) * mean;
    mvar[c] = momentum * mvar[c] + (1.f - momentum) * (float)unb;
  } else {
    mean = mmean[c];
    var = mvar[c];
  }
  float is = rsqrtf(var + eps);
  is = is * (1.5f - 0.5f * (var + eps) * is * is);   // one Newton step: full fp32 accuracy
  float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (save_mean) save_mean[c] = mean;
  if (save_invstd) save_invstd[c] = is;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y,
                                                            int ldy, int C, long long npix,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift,
                                                            double* __restrict__ out_stats, int out_sq_off,
                                                            const T* __restrict__ x2, int ldx2, int split,
                                                            const uint8_t* __restrict__ drop_bits, float p_drop) {
  // (drop_bits, p_drop): the input is dropout(x) of an unmaterialised Dropout layer, keep mask as stored by the statistics
  // pass (see bn_reduce_kernel); p_drop = 0: none
  // (x2, ldx2, split): channels [split, C) are read from a second tensor -- a two-input concatenate whose halves live in
  // dense tensors of their own instead of one interleaved buffer (plan.py "split concat"); split = C: one source
  B2U_PDL_PROLOGUE();
  extern __shared__ float sst[];          // [2*C] block partials of the optional output statistics
  float o1[8], o2[8];
  if (out_stats != nullptr) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sst[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { o1[k] = 0.f; o2[k] = 0.f; }
    __syncthreads();
  }
  float sc[8], sh[8];
  {
    const int g0 = (threadIdx.x % (C >> 3)) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[g0 + k]; sh[k] = shift[g0 + k]; }
  }
  PIXEL_LANE_LOOP(C, npix) {
    float v[8];
    load8<T>(g * 8 < split ? x + p * ldx + g * 8 : x2 + p * ldx2 + (g * 8 - split), v);
    if (p_drop > 0.f) {
      float f[8];
      bits_factors(drop_bits[((uint64_t)p * C + g * 8) >> 3], 1.f / (1.f - p_drop), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= f[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
    store8<T>(y + p * ldy + g * 8, v);
    if (out_stats != nullptr) {
      float r[8];
      load8<T>(y + p * ldy + g * 8, r);      // statistics of the value as stored (fp16 rounding included)
#pragma unroll
      for (int k = 0; k < 8; ++k) { o1[k] += r[k]; o2[k] = fmaf(r[k], r[k], o2[k]); }
    }
  }
  if (out_stats != nullptr) {
    if (threadIdx.x / (C >> 3) < kThreads / (C >> 3)) {
      group_add8(sst, o1, C >> 3, threadIdx.x % (C >> 3));
      group_add8(sst + C, o2, C >> 3, threadIdx.x % (C >> 3));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      atomicAdd(&out_stats[i], (double)sst[i]);
      atomicAdd(&out_stats[out_sq_off + i], (double)sst[C + i]);
    }
  }
}

// kTwo: two pixels per trip with every load issued before the first use.  Measured on B200 at 512^2 x 32 channels, batch 8
// (tools/bn_bwd_probe.py): the plain variant is fastest with one pixel per trip and its higher occupancy (0.074 against
// 0.085 ms), the variants with column sums / an activation mask / dropout bits with two (0.097 -> 0.093, 0.133 -> 0.101).
template <typename T, bool kColsum, bool kTwo>
__global__ void __launch_bounds__(kThreads, kTwo ? 2 : 4) bn_bwd_apply_kernel(
    const T* __restrict__ dy, int lddy, const T* __restrict__ x, int ldx, T* __restrict__ dx, int lddx, int C,
    long long npix, long long count, const float* __restrict__ gamma, const float* __restrict__ mean,
    const float* __restrict__ invstd, const double* __restrict__ sums, float* __restrict__ dgamma,
    float* __restrict__ dbeta, const T* __restrict__ mask, int ldmask, int mask_act, float* __restrict__ colsum,
    const T* __restrict__ x2, int ldx2, T* __restrict__ dx2, int lddx2, int split,
    const uint8_t* __restrict__ drop_bits, float p_drop) {
  B2U_PDL_PROLOGUE();
  // (drop_bits, p_drop): the BN input is dropout(x) of an unmaterialised Dropout layer: x * f enters the BN backward and the
  // gradient written is the one of x itself (times f: the dropout backward), p_drop = 0: none
  // (x2, dx2, split): channels [split, C) of the BN input / its gradient live in a second tensor (split concat, see
  // bn_apply_kernel); split = C: one tensor
  // colsum (optional): per-channel sums of the dx values written = bias gradient of the conv that feeds this BN
  extern __shared__ float scs[];          // [C] block partials of colsum
  float cs[kColsum ? 8 : 1];
#pragma unroll
  for (int k = 0; k < (kColsum ? 8 : 1); ++k) cs[k] = 0.f;
  if (kColsum) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) scs[i] = 0.f;
    __syncthreads();
  }
  const float inv_n = 1.f / (float)count;
  if (blockIdx.x == 0 && dgamma != nullptr) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] += (float)sums[c];
      dgamma[c] += (float)sums[C + c];
    }
  }
  // per-channel constants in registers: dx = a*dy + b*x + c0  with
  //   a = gamma*invstd, b = -a*invstd*s2, c0 = -a*s1 + a*invstd*mean*s2
  float ca[8], cb[8], cc[8];
  {
    const int g0 = (threadIdx.x % (C >> 3)) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g0 + k;
      const float is = invstd[c], mu = mean[c];
      const float s1 = (float)(sums[c] * (double)inv_n), s2 = (float)(sums[C + c] * (double)inv_n);
      const float a = gamma[c] * is;
      ca[k] = a;
      cb[k] = -a * is * s2;
      cc[k] = -a * s1 + a * is * mu * s2;
    }
  }
  // the activation mask of a BN whose input is the activated conv output itself (mask == x) is taken from the registers
  // that already hold x
  const int cg = C >> 3;
  const int lanes = kThreads / cg;
  const int g = threadIdx.x % cg, lane_ = threadIdx.x / cg;
  const bool first = g * 8 < split;
  const bool mask_is_x = mask != nullptr && mask == x && ldmask == ldx && split == C;
  const float dsc = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  if (lane_ < lanes) {
    const long long stride = (long long)gridDim.x * lanes;
    constexpr int kU = kTwo ? 2 : 1;
    for (long long p0 = (long long)blockIdx.x * lanes + lane_; p0 < npix; p0 += kU * stride) {
      const bool two = kTwo && p0 + stride < npix;
      Pack8<T> pd[kU], px[kU], pm[kU];                    // packed as loaded: 12 registers in flight per pixel (fp16)
      unsigned db[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        db[u] = 0xffu;
        if (u == 0 || two) {
          const long long p = p0 + u * stride;
          pd[u].load(dy + p * lddy + g * 8);
          px[u].load(first ? x + p * ldx + g * 8 : x2 + p * ldx2 + (g * 8 - split));
          if (mask != nullptr && !mask_is_x) pm[u].load(mask + p * ldmask + g * 8);
          if (p_drop > 0.f) db[u] = drop_bits[((uint64_t)p * C + g * 8) >> 3];
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (u == 1 && !two) continue;
        const long long p = p0 + u * stride;
        float o[8], d[8], xv[8];
        pd[u].get(d);
        px[u].get(xv);
        if (p_drop > 0.f) {
          float f[8];
          bits_factors(db[u], dsc, f);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = f[k] * fmaf(ca[k], d[k], fmaf(cb[k], xv[k] * f[k], cc[k]));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = fmaf(ca[k], d[k], fmaf(cb[k], xv[k], cc[k]));
        }
        if (mask != nullptr) {
          float mv[8];
          if (!mask_is_x) pm[u].get(mv);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] *= act_bwd_from_y(mask_is_x ? xv[k] : mv[k], mask_act);
        }
        store8<T>(first ? dx + p * lddx + g * 8 : dx2 + p * lddx2 + (g * 8 - split), o);
        if (kColsum) {
#pragma unroll
          for (int k = 0; k < 8; ++k) cs[kColsum ? k : 0] += o[k];
        }
      }
    }
  }
  if (kColsum) {
    if (threadIdx.x / (C >> 3) < kThreads / (C >> 3)) {
      float c8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) c8[k] = cs[kColsum ? k : 0];
      group_add8(scs, c8, C >> 3, threadIdx.x % (C >> 3));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&colsum[i], scs[i]);
  }
}

// The plain fp16 case (decoder BatchNorms: no mask, no column sums, one tensor) with the loads decoupled from the
// registers: every thread keeps kAD (dy, x) pairs of 16-byte cp.async copies in flight into its own shared-memory slots
// (4 blocks x 256 threads x kAD x 32 B = 128 KB of loads in flight per SM against 32 KB of the register version, which
// ran at 5.4 TB/s); a thread only ever reads the slots it filled itself, so cp.async.wait_group is all the
// synchronisation there is.
constexpr int kAD = 4;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
// BN apply, same scheme; it has ONE input stream, so the ring is kAA = 8 deep for the same 128 KB in flight per SM;
// channels [split, C) from a second tensor
constexpr int kAA = 8;
__global__ void __launch_bounds__(kThreads, 4) bn_apply_async_kernel(const __half* __restrict__ x, int ldx,
                                                                     __half* __restrict__ y, int ldy, int C, long long npix,
                                                                     const float* __restrict__ scale,
                                                                     const float* __restrict__ shift,
                                                                     const __half* __restrict__ x2, int ldx2, int split,
                                                                     const uint8_t* __restrict__ drop_bits, float p_drop) {
  // (drop_bits, p_drop): the input is dropout(x) of an unmaterialised Dropout layer (keep bits stored by the statistics pass)
  B2U_PDL_PROLOGUE();
  extern __shared__ __align__(16) uint4 ring[];            // [kAA][kThreads]
  const int cg = C >> 3;
  const int lanes = kThreads / cg;
  const int g = threadIdx.x % cg, lane_ = threadIdx.x / cg;
  if (lane_ >= lanes) return;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = scale[g * 8 + k]; sh[k] = shift[g * 8 + k]; }
  const __half* src = g * 8 < split ? x + g * 8 : x2 + (g * 8 - split);
  const long long lds = g * 8 < split ? ldx : ldx2;
  const float dsc = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const long long stride = (long long)gridDim.x * lanes;
  uint4* my = ring + threadIdx.x;
  long long pl = (long long)blockIdx.x * lanes + lane_;
#pragma unroll
  for (int d = 0; d < kAA; ++d) {
    if (pl < npix) cp_async16(my + d * kThreads, src + pl * lds);
    asm volatile("cp.async.commit_group;" ::: "memory");
    pl += stride;
  }
  int slot = 0;
  for (long long p = (long long)blockIdx.x * lanes + lane_; p < npix; p += stride) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kAA - 1) : "memory");
    const uint4 ux = my[slot * kThreads];
    if (pl < npix) cp_async16(my + slot * kThreads, src + pl * lds);
    asm volatile("cp.async.commit_group;" ::: "memory");
    pl += stride;
    float v[8];
    unpack8h(ux, v);
    if (p_drop > 0.f) {
      float f[8];
      bits_factors(drop_bits[((uint64_t)p * C + g * 8) >> 3], dsc, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= f[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
    store8<__half>(y + p * ldy + g * 8, v);
    if (++slot == kAA) slot = 0;
  }
}

// kColsum / mask_act: the encoder BatchNorms -- column sums of the gradient written (bias gradient of the producing conv) and
// its activation derivative taken from x itself (mask == x).
template <bool kColsum>
__global__ void __launch_bounds__(kThreads, kColsum ? 3 : 4) bn_bwd_apply_async_kernel(
    const __half* __restrict__ dy, int lddy, const __half* __restrict__ x, int ldx, __half* __restrict__ dx, int lddx, int C,
    long long npix, long long count, const float* __restrict__ gamma, const float* __restrict__ mean,
    const float* __restrict__ invstd, const double* __restrict__ sums, float* __restrict__ dgamma,
    float* __restrict__ dbeta, int mask_act, float* __restrict__ colsum, const __half* __restrict__ x2, int ldx2,
    __half* __restrict__ dx2, int lddx2, int split, const uint8_t* __restrict__ drop_bits, float p_drop) {
  // (x2, dx2, split): channels [split, C) of the BN input / its gradient live in a second tensor (split concatenate)
  // (drop_bits, p_drop): the BN input is dropout(x) of an unmaterialised Dropout layer (see bn_bwd_apply_kernel)
  B2U_PDL_PROLOGUE();
  extern __shared__ __align__(16) uint4 ring[];            // [kAD][2][kThreads], then [C] floats of column-sum partials
  float* scs = reinterpret_cast<float*>(ring + kAD * 2 * kThreads);
  float cs[kColsum ? 8 : 1];
#pragma unroll
  for (int k = 0; k < (kColsum ? 8 : 1); ++k) cs[k] = 0.f;
  if (kColsum) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) scs[i] = 0.f;
    __syncthreads();
  }
  const float inv_n = 1.f / (float)count;
  if (blockIdx.x == 0 && dgamma != nullptr) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] += (float)sums[c];
      dgamma[c] += (float)sums[C + c];
    }
  }
  float ca[8], cb[8], cc[8];
  {
    const int g0 = (threadIdx.x % (C >> 3)) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g0 + k;
      const float is = invstd[c], mu = mean[c];
      const float s1 = (float)(sums[c] * (double)inv_n), s2 = (float)(sums[C + c] * (double)inv_n);
      const float a = gamma[c] * is;
      ca[k] = a;
      cb[k] = -a * is * s2;
      cc[k] = -a * s1 + a * is * mu * s2;
    }
  }
  const int cg = C >> 3;
  const int lanes = kThreads / cg;
  const int g = threadIdx.x % cg, lane_ = threadIdx.x / cg;
  if (lane_ < lanes) {
  const long long stride = (long long)gridDim.x * lanes;
  uint4* my = ring + threadIdx.x;
  const bool first = g * 8 < split;
  const __half* xs = first ? x + g * 8 : x2 + (g * 8 - split);       // this thread's channel group: source / destination
  __half* ds = first ? dx + g * 8 : dx2 + (g * 8 - split);
  const long long lxs = first ? ldx : ldx2, lds = first ? lddx : lddx2;
  const float dsc = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  long long pl = (long long)blockIdx.x * lanes + lane_;     // next pixel to load
#pragma unroll
  for (int d = 0; d < kAD; ++d) {
    if (pl < npix) {
      cp_async16(my + (d * 2 + 0) * kThreads, dy + pl * lddy + g * 8);
      cp_async16(my + (d * 2 + 1) * kThreads, xs + pl * lxs);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pl += stride;
  }
  int slot = 0;
  for (long long p = (long long)blockIdx.x * lanes + lane_; p < npix; p += stride) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kAD - 1) : "memory");
    const uint4 ud = my[(slot * 2 + 0) * kThreads], ux = my[(slot * 2 + 1) * kThreads];
    if (pl < npix) {                                        // refill the slot just read
      cp_async16(my + (slot * 2 + 0) * kThreads, dy + pl * lddy + g * 8);
      cp_async16(my + (slot * 2 + 1) * kThreads, xs + pl * lxs);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pl += stride;
    float d[8], xv[8], o[8];
    unpack8h(ud, d);
    unpack8h(ux, xv);
    if (p_drop > 0.f) {
      float f[8];
      bits_factors(drop_bits[((uint64_t)p * C + g * 8) >> 3], dsc, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = f[k] * fmaf(ca[k], d[k], fmaf(cb[k], xv[k] * f[k], cc[k]));
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(ca[k], d[k], fmaf(cb[k], xv[k], cc[k]));
    }
    if (mask_act != B2U_ACT_NONE) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] *= act_bwd_from_y(xv[k], mask_act);
    }
    store8<__half>(ds + p * lds, o);
    if (kColsum) {
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[kColsum ? k : 0] += o[k];
    }
    if (++slot == kAD) slot = 0;
  }
  }
  if (kColsum) {
    if (lane_ < lanes) {
      float c8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) c8[k] = cs[kColsum ? k : 0];
      group_add8(scs, c8, cg, g);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&colsum[i], scs[i]);
  }
}

// ------------------------------------------------------------------------------------------
// 2x2 max-pool (+ optional dropout on the pooled output)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void keep_factors(float p, uint64_t e0, const b2u_step_state* st, int op_id,
                                             float f[8]) {
  // e0 is the logical index of the first of 8 consecutive elements (multiple of 8)
  const uint32_t thr = dropout_threshold(p);
  const float sc = 1.f / (1.f - p);
  uint4 a = dropout_words(e0 >> 2, st->seed, st->step, (uint32_t)op_id);
  uint4 b = dropout_words((e0 >> 2) + 1, st->seed, st->step, (uint32_t)op_id);
  f[0] = a.x >= thr ? sc : 0.f; f[1] = a.y >= thr ? sc : 0.f; f[2] = a.z >= thr ? sc : 0.f; f[3] = a.w >= thr ? sc : 0.f;
  f[4] = b.x >= thr ? sc : 0.f; f[5] = b.y >= thr ? sc : 0.f; f[6] = b.z >= thr ? sc : 0.f; f[7] = b.w >= thr ? sc : 0.f;
}

template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<__half>(float v) { return __half2float(__float2half_rn(v)); }

// BN apply + 2x2 max-pool (+ dropout) in one pass: thread = (output pixel, 8-channel group) owns the 2x2 window
template <typename T>
__global__ void __launch_bounds__(kThreads) bn_apply_pool_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y,
                                                                 int ldy, int C, int N, int H, int W,
                                                                 const float* __restrict__ scale,
                                                                 const float* __restrict__ shift,
                                                                 double* __restrict__ out_stats, int out_sq_off,
                                                                 T* __restrict__ yp, int ldp, float p_drop, int op_id,
                                                                 const b2u_step_state* __restrict__ st) {
  B2U_PDL_PROLOGUE();
  extern __shared__ float sst[];          // [2*C] block partials of the optional output statistics
  float o1[8], o2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { o1[k] = 0.f; o2[k] = 0.f; }
  if (out_stats != nullptr) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sst[i] = 0.f;
    __syncthreads();
  }
  float sc[8], sh[8];
  {
    const int g0 = (threadIdx.x % (C >> 3)) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[g0 + k]; sh[k] = shift[g0 + k]; }
  }
  const int Ho = H >> 1, Wo = W >> 1;
  const long long nopix = (long long)N * Ho * Wo;
  PIXEL_LANE_LOOP(C, nopix) {
    const unsigned opu = (unsigned)p;
    const int wo = (int)(opu % (unsigned)Wo);
    const unsigned t = opu / (unsigned)Wo;
    const int ho = (int)(t % (unsigned)Ho);
    const int n = (int)(t / (unsigned)Ho);
    const long long ip = ((long long)n * H + 2 * ho) * W + 2 * wo;
    Pack8<T> xw[4];
    xw[0].load(x + ip * ldx + g * 8);
    xw[1].load(x + (ip + 1) * ldx + g * 8);
    xw[2].load(x + (ip + W) * ldx + g * 8);
    xw[3].load(x + (ip + W + 1) * ldx + g * 8);
    float m[8];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      float v[8];
      xw[pos].get(v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[k] = round_to<T>(fmaf(v[k], sc[k], sh[k]));              // the value as stored
        m[k] = pos == 0 ? v[k] : fmaxf(m[k], v[k]);
        o1[k] += v[k];
        o2[k] = fmaf(v[k], v[k], o2[k]);
      }
      store8<T>(y + (ip + (pos >> 1) * (long long)W + (pos & 1)) * ldy + g * 8, v);
    }
    if (p_drop > 0.f) {
      float f[8];
      keep_factors(p_drop, (uint64_t)p * C + g * 8, st, op_id, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] *= f[k];
    }
    store8<T>(yp + p * ldp + g * 8, m);
  }
  if (out_stats != nullptr) {
    if (threadIdx.x / (C >> 3) < kThreads / (C >> 3)) {
      group_add8(sst, o1, C >> 3, threadIdx.x % (C >> 3));
      group_add8(sst + C, o2, C >> 3, threadIdx.x % (C >> 3));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      atomicAdd(&out_stats[i], (double)sst[i]);
      atomicAdd(&out_stats[out_sq_off + i], (double)sst[C + i]);
    }
  }
}

// fp16 cp.async version (see bn_bwd_apply_async_kernel): the four window loads of kPD output pixels in flight per thread
constexpr int kPD = 4;
__global__ void __launch_bounds__(kThreads, 2) bn_apply_pool_async_kernel(
    const __half* __restrict__ x, int ldx, __half* __restrict__ y, int ldy, int C, int N, int H, int W,
    const float* __restrict__ scale, const float* __restrict__ shift, double* __restrict__ out_stats, int out_sq_off,
    __half* __restrict__ yp, int ldp, float p_drop, int op_id, const b2u_step_state* __restrict__ st) {
  B2U_PDL_PROLOGUE();
  extern __shared__ __align__(16) uint4 ring[];            // [kPD][4][kThreads], then [2*C] floats of statistics partials
  float* sst = reinterpret_cast<float*>(ring + kPD * 4 * kThreads);
  float o1[8], o2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { o1[k] = 0.f; o2[k] = 0.f; }
  if (out_stats != nullptr) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sst[i] = 0.f;
    __syncthreads();
  }
  const int cg = C >> 3;
  const int lanes = kThreads / cg;
  const int g = threadIdx.x % cg, lane_ = threadIdx.x / cg;
  if (lane_ < lanes) {
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[g * 8 + k]; sh[k] = shift[g * 8 + k]; }
    const int Ho = H >> 1, Wo = W >> 1;
    const long long nopix = (long long)N * Ho * Wo;
    const long long stride = (long long)gridDim.x * lanes;
    uint4* my = ring + threadIdx.x;
    auto in_pix = [&](long long op) {                      // first input pixel of output pixel op's window
      const unsigned opu = (unsigned)op;
      const int wo = (int)(opu % (unsigned)Wo);
      const unsigned t = opu / (unsigned)Wo;
      const int ho = (int)(t % (unsigned)Ho);
      const int n = (int)(t / (unsigned)Ho);
      return ((long long)n * H + 2 * ho) * W + 2 * wo;
    };
    auto issue = [&](int slot, long long op) {
      const long long ip = in_pix(op);
      const __half* b = x + ip * ldx + g * 8;
      cp_async16(my + (slot * 4 + 0) * kThreads, b);
      cp_async16(my + (slot * 4 + 1) * kThreads, b + ldx);
      cp_async16(my + (slot * 4 + 2) * kThreads, b + (long long)W * ldx);
      cp_async16(my + (slot * 4 + 3) * kThreads, b + (long long)W * ldx + ldx);
    };
    long long pl = (long long)blockIdx.x * lanes + lane_;
#pragma unroll
    for (int d = 0; d < kPD; ++d) {
      if (pl < nopix) issue(d, pl);
      asm volatile("cp.async.commit_group;" ::: "memory");
      pl += stride;
    }
    int slot = 0;
    for (long long p = (long long)blockIdx.x * lanes + lane_; p < nopix; p += stride) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kPD - 1) : "memory");
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) u[q] = my[(slot * 4 + q) * kThreads];
      if (pl < nopix) issue(slot, pl);
      asm volatile("cp.async.commit_group;" ::: "memory");
      pl += stride;
      const long long ip = in_pix(p);
      float m[8];
#pragma unroll
      for (int pos = 0; pos < 4; ++pos) {
        float v[8];
        unpack8h(u[pos], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = round_to<__half>(fmaf(v[k], sc[k], sh[k]));        // the value as stored
          m[k] = pos == 0 ? v[k] : fmaxf(m[k], v[k]);
          o1[k] += v[k];
          o2[k] = fmaf(v[k], v[k], o2[k]);
        }
        store8<__half>(y + (ip + (pos >> 1) * (long long)W + (pos & 1)) * ldy + g * 8, v);
      }
      if (p_drop > 0.f) {
        float f[8];
        keep_factors(p_drop, (uint64_t)p * C + g * 8, st, op_id, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] *= f[k];
      }
      store8<__half>(yp + p * ldp + g * 8, m);
      if (++slot == kPD) slot = 0;
    }
  }
  if (out_stats != nullptr) {
    if (lane_ < lanes) {
      group_add8(sst, o1, cg, g);
      group_add8(sst + C, o2, cg, g);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      atomicAdd(&out_stats[i], (double)sst[i]);
      atomicAdd(&out_stats[out_sq_off + i], (double)sst[C + i]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) maxpool_fwd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y,
                                                               int ldy, int C, int N, int H, int W, float p_drop,
                                                               int op_id, const b2u_step_state* __restrict__ st) {
  B2U_PDL_PROLOGUE();
  const int Ho = H >> 1, Wo = W >> 1;
  const long long nopix = (long long)N * Ho * Wo;            // < 2^31 for any tensor that fits in HBM here
  PIXEL_LANE_LOOP(C, nopix) {
    const long long op = p;
    const unsigned opu = (unsigned)p;
    const int wo = (int)(opu % (unsigned)Wo);
    const unsigned t = opu / (unsigned)Wo;
    const int ho = (int)(t % (unsigned)Ho);
    const int n = (int)(t / (unsigned)Ho);
    const T* base = x + (((long long)n * H + 2 * ho) * W + 2 * wo) * ldx + g * 8;
    float a[8], b[8], c[8], d[8], m[8];
    load8<T>(base, a);
    load8<T>(base + ldx, b);
    load8<T>(base + (long long)W * ldx, c);
    load8<T>(base + (long long)W * ldx + ldx, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
    if (p_drop > 0.f) {
      float f[8];
      keep_factors(p_drop, (uint64_t)op * C + g * 8, st, op_id, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] *= f[k];
    }
    store8<T>(y + op * ldy + g * 8, m);
  }
}

// Register budget matters here: the four window values stay PACKED (as loaded) and are unpacked one window
// position at a time, so three 256-thread blocks fit an SM (the all-unpacked version needed 158 registers:
// one block per SM, 12 % occupancy, a third of the HBM rate).
template <typename T>
__global__ void __launch_bounds__(kThreads, 2) maxpool_bwd_kernel(const T* __restrict__ x, int ldx,
                                                                  const T* __restrict__ dy, int lddy,
                                                                  T* __restrict__ dx, int lddx, int C, int N, int H,
                                                                  int W, float p_drop, int op_id,
                                                                  const b2u_step_state* __restrict__ st,
                                                                  int accumulate, double* __restrict__ bn_sums,
                                                                  const float* __restrict__ bn_gamma,
                                                                  const float* __restrict__ bn_beta) {
  B2U_PDL_PROLOGUE();
  extern __shared__ float sbn[];          // [2*C] block partials of the fused BN-backward statistics
  float t1[8], t2[8], bt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { t1[k] = 0.f; t2[k] = 0.f; bt[k] = 0.f; }
  if (bn_sums != nullptr) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sbn[i] = 0.f;
    const int g0 = (threadIdx.x % (C >> 3)) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) bt[k] = bn_beta[g0 + k];
    __syncthreads();
  }
  const int Ho = H >> 1, Wo = W >> 1;
  const long long nopix = (long long)N * Ho * Wo;            // < 2^31 for any tensor that fits in HBM here
  PIXEL_LANE_LOOP(C, nopix) {
    const long long op = p;
    const unsigned opu = (unsigned)p;
    const int wo = (int)(opu % (unsigned)Wo);
    const unsigned t = opu / (unsigned)Wo;
    const int ho = (int)(t % (unsigned)Ho);
    const int n = (int)(t / (unsigned)Ho);
    const long long ip = ((long long)n * H + 2 * ho) * W + 2 * wo;
    const T* base = x + ip * ldx + g * 8;
    T* ob_ = dx + ip * lddx + g * 8;
    // all loads of the window first (packed), then the arithmetic
    Pack8<T> xw[4], ew[4], gp;
    xw[0].load(base);
    xw[1].load(base + ldx);
    xw[2].load(base + (long long)W * ldx);
    xw[3].load(base + (long long)W * ldx + ldx);
    gp.load(dy + op * lddy + g * 8);
    if (accumulate) {
      ew[0].load(ob_);
      ew[1].load(ob_ + lddx);
      ew[2].load(ob_ + (long long)W * lddx);
      ew[3].load(ob_ + (long long)W * lddx + lddx);
    }
    float gy[8];
    gp.get(gy);
    if (p_drop > 0.f) {
      float f[8];
      keep_factors(p_drop, (uint64_t)op * C + g * 8, st, op_id, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) gy[k] *= f[k];
    }
    // first maximum in window order (0,0),(0,1),(1,0),(1,1) takes the gradient (TF/torch tie rule)
    int sel[8];
    {
      float a[8], b[8], m[8];
      xw[0].get(a);
      xw[1].get(b);
#pragma unroll
      for (int k = 0; k < 8; ++k) { sel[k] = b[k] > a[k] ? 1 : 0; m[k] = fmaxf(a[k], b[k]); }
      xw[2].get(a);
      xw[3].get(b);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (a[k] > m[k]) { sel[k] = 2; m[k] = a[k]; }
        if (b[k] > m[k]) { sel[k] = 3; }
      }
    }
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = sel[k] == pos ? gy[k] : 0.f;
      if (accumulate) {
        float e[8];
        ew[pos].get(e);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += e[k];
      }
      store8<T>(ob_ + (pos >> 1) * (long long)W * lddx + (pos & 1) * lddx, o);
      if (bn_sums != nullptr) {
        float xv[8];
        xw[pos].get(xv);
#pragma unroll
        for (int k = 0; k < 8; ++k) { t1[k] += o[k]; t2[k] = fmaf(o[k], xv[k] - bt[k], t2[k]); }
      }
    }
  }
  if (bn_sums != nullptr) {
    // sum dx * xhat with xhat = (x - beta) / gamma: the division by gamma is applied once per thread
    const int g0 = (threadIdx.x % (C >> 3)) * 8;
    if (threadIdx.x / (C >> 3) < kThreads / (C >> 3)) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gm = bn_gamma[g0 + k];
        t2[k] *= fabsf(gm) > 1e-12f ? 1.f / gm : 0.f;
      }
      group_add8(sbn, t1, C >> 3, threadIdx.x % (C >> 3));
      group_add8(sbn + C, t2, C >> 3, threadIdx.x % (C >> 3));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&bn_sums[i], (double)sbn[i]);
  }
}

// fp16 cp.async version: the nine 16-byte loads of an output pixel (four window values, the pooled gradient, the four
// gradients accumulated into) for kMD output pixels in flight per thread
constexpr int kMD = 2;
__global__ void __launch_bounds__(kThreads, 2) maxpool_bwd_async_kernel(
    const __half* __restrict__ x, int ldx, const __half* __restrict__ dy, int lddy, __half* __restrict__ dx, int lddx, int C,
    int N, int H, int W, float p_drop, int op_id, const b2u_step_state* __restrict__ st, int accumulate,
    double* __restrict__ bn_sums, const float* __restrict__ bn_gamma, const float* __restrict__ bn_beta) {
  B2U_PDL_PROLOGUE();
  extern __shared__ __align__(16) uint4 ring[];            // [kMD][9][kThreads], then [2*C] floats of BN-statistics partials
  float* sbn = reinterpret_cast<float*>(ring + kMD * 9 * kThreads);
  float t1[8], t2[8], bt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { t1[k] = 0.f; t2[k] = 0.f; bt[k] = 0.f; }
  const int cg = C >> 3;
  const int lanes = kThreads / cg;
  const int g = threadIdx.x % cg, lane_ = threadIdx.x / cg;
  if (bn_sums != nullptr) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sbn[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) bt[k] = bn_beta[g * 8 + k];
    __syncthreads();
  }
  if (lane_ < lanes) {
    const int Ho = H >> 1, Wo = W >> 1;
    const long long nopix = (long long)N * Ho * Wo;
    const long long stride = (long long)gridDim.x * lanes;
    uint4* my = ring + threadIdx.x;
    auto in_pix = [&](long long op) {
      const unsigned opu = (unsigned)op;
      const int wo = (int)(opu % (unsigned)Wo);
      const unsigned t = opu / (unsigned)Wo;
      const int ho = (int)(t % (unsigned)Ho);
      const int n = (int)(t / (unsigned)Ho);
      return ((long long)n * H + 2 * ho) * W + 2 * wo;
    };
    auto issue = [&](int slot, long long op) {
      const long long ip = in_pix(op);
      const __half* b = x + ip * ldx + g * 8;
      uint4* s0 = my + (slot * 9) * kThreads;
      cp_async16(s0 + 0 * kThreads, b);
      cp_async16(s0 + 1 * kThreads, b + ldx);
      cp_async16(s0 + 2 * kThreads, b + (long long)W * ldx);
      cp_async16(s0 + 3 * kThreads, b + (long long)W * ldx + ldx);
      cp_async16(s0 + 4 * kThreads, dy + op * lddy + g * 8);
      if (accumulate) {
        const __half* e = dx + ip * lddx + g * 8;
        cp_async16(s0 + 5 * kThreads, e);
        cp_async16(s0 + 6 * kThreads, e + lddx);
        cp_async16(s0 + 7 * kThreads, e + (long long)W * lddx);
        cp_async16(s0 + 8 * kThreads, e + (long long)W * lddx + lddx);
      }
    };
    long long pl = (long long)blockIdx.x * lanes + lane_;
#pragma unroll
    for (int d = 0; d < kMD; ++d) {
      if (pl < nopix) issue(d, pl);
      asm volatile("cp.async.commit_group;" ::: "memory");
      pl += stride;
    }
    int slot = 0;
    for (long long p = (long long)blockIdx.x * lanes + lane_; p < nopix; p += stride) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kMD - 1) : "memory");
      uint4 xw[4], ew[4], gp;
      const uint4* s0 = my + (slot * 9) * kThreads;
#pragma unroll
      for (int q = 0; q < 4; ++q) xw[q] = s0[q * kThreads];
      gp = s0[4 * kThreads];
      if (accumulate) {
#pragma unroll
        for (int q = 0; q < 4; ++q) ew[q] = s0[(5 + q) * kThreads];
      }
      if (pl < nopix) issue(slot, pl);
      asm volatile("cp.async.commit_group;" ::: "memory");
      pl += stride;
      const long long ip = in_pix(p);
      __half* ob_ = dx + ip * lddx + g * 8;
      float gy[8];
      unpack8h(gp, gy);
      if (p_drop > 0.f) {
        float f[8];
        keep_factors(p_drop, (uint64_t)p * C + g * 8, st, op_id, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) gy[k] *= f[k];
      }
      // first maximum in window order (0,0),(0,1),(1,0),(1,1) takes the gradient (TF/torch tie rule)
      int sel[8];
      {
        float a[8], b[8], m[8];
        unpack8h(xw[0], a);
        unpack8h(xw[1], b);
#pragma unroll
        for (int k = 0; k < 8; ++k) { sel[k] = b[k] > a[k] ? 1 : 0; m[k] = fmaxf(a[k], b[k]); }
        unpack8h(xw[2], a);
        unpack8h(xw[3], b);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (a[k] > m[k]) { sel[k] = 2; m[k] = a[k]; }
          if (b[k] > m[k]) { sel[k] = 3; }
        }
      }
#pragma unroll
      for (int pos = 0; pos < 4; ++pos) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = sel[k] == pos ? gy[k] : 0.f;
        if (accumulate) {
          float e[8];
          unpack8h(ew[pos], e);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += e[k];
        }
        store8<__half>(ob_ + (pos >> 1) * (long long)W * lddx + (pos & 1) * lddx, o);
        if (bn_sums != nullptr) {
          float xv[8];
          unpack8h(xw[pos], xv);
#pragma unroll
          for (int k = 0; k < 8; ++k) { t1[k] += o[k]; t2[k] = fmaf(o[k], xv[k] - bt[k], t2[k]); }
        }
      }
      if (++slot == kMD) slot = 0;
    }
  }
  if (bn_sums != nullptr) {
    if (lane_ < lanes) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gm = bn_gamma[g * 8 + k];
        t2[k] *= fabsf(gm) > 1e-12f ? 1.f / gm : 0.f;
      }
      group_add8(sbn, t1, cg, g);
      group_add8(sbn + C, t2, cg, g);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&bn_sums[i], (double)sbn[i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) dropout_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y,
                                                           int ldy, int C, long long npix, float prob, int op_id,
                                                           const b2u_step_state* __restrict__ st,
                                                           const T* __restrict__ mask, int ldmask, int mask_act) {
  B2U_PDL_PROLOGUE();
  PIXEL_LANE_LOOP(C, npix) {
    const long long px = p;
    float v[8], f[8];
    load8<T>(x + px * ldx + g * 8, v);
    keep_factors(prob, (uint64_t)px * C + g * 8, st, op_id, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= f[k];
    if (mask != nullptr) {
      float mv[8];
      load8<T>(mask + px * ldmask + g * 8, mv);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= act_bwd_from_y(mv[k], mask_act);
    }
    store8<T>(y + px * ldy + g * 8, v);
  }
}

// packed 1-bit ReLU masks (bit index pix*C + c, little-endian within bytes): one byte per thread and pixel
template <typename T>
__global__ void __launch_bounds__(kThreads) relu_bits_kernel(const T* __restrict__ y, int ldy, int C, long long npix,
                                                             uint8_t* __restrict__ bits) {
  B2U_PDL_PROLOGUE();
  PIXEL_LANE_LOOP(C, npix) {
    float v[8];
    load8<T>(y + p * ldy + g * 8, v);
    unsigned b = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) b |= (v[k] > 0.f ? 1u : 0u) << k;
    bits[((p * C) >> 3) + g] = (uint8_t)b;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) apply_bits_kernel(T* __restrict__ dx, int lddx, int C, long long npix,
                                                              const uint8_t* __restrict__ bits) {
  B2U_PDL_PROLOGUE();
  PIXEL_LANE_LOOP(C, npix) {
    float v[8];
    load8<T>(dx + p * lddx + g * 8, v);
    const unsigned b = bits[((p * C) >> 3) + g];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ((b >> k) & 1u) ? v[k] : 0.f;
    store8<T>(dx + p * lddx + g * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) copy_slice_kernel(const T* __restrict__ s, int lds, T* __restrict__ d,
                                                              int ldd, int C, long long npix, int accumulate) {
  B2U_PDL_PROLOGUE();
  PIXEL_LANE_LOOP(C, npix) {
    const long long px = p;
    float v[8];
    load8<T>(s + px * lds + g * 8, v);
    if (accumulate) {
      float e[8];
      load8<T>(d + px * ldd + g * 8, e);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += e[k];
    }
    store8<T>(d + px * ldd + g * 8, v);
  }
}

// fp16 cp.async version (the multi-consumer concat copies of U-Net++): source (and, accumulating, destination) vectors kAA
// deep in flight per thread, as in bn_apply_async_kernel
__global__ void __launch_bounds__(kThreads, 4) copy_slice_async_kernel(const __half* __restrict__ s, int lds,
                                                                       __half* __restrict__ d, int ldd, int C, long long npix,
                                                                       int accumulate) {
  B2U_PDL_PROLOGUE();
  extern __shared__ __align__(16) uint4 ring[];            // [kAA / 2][2][kThreads]
  constexpr int kD = kAA / 2;
  const int cg = C >> 3;
  const int lanes = kThreads / cg;
  const int g = threadIdx.x % cg, lane_ = threadIdx.x / cg;
  if (lane_ >= lanes) return;
  const long long stride = (long long)gridDim.x * lanes;
  uint4* my = ring + threadIdx.x;
  long long pl = (long long)blockIdx.x * lanes + lane_;
#pragma unroll
  for (int q = 0; q < kD; ++q) {
    if (pl < npix) {
      cp_async16(my + (q * 2 + 0) * kThreads, s + pl * lds + g * 8);
      if (accumulate) cp_async16(my + (q * 2 + 1) * kThreads, d + pl * ldd + g * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pl += stride;
  }
  int slot = 0;
  for (long long p = (long long)blockIdx.x * lanes + lane_; p < npix; p += stride) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kD - 1) : "memory");
    const uint4 us = my[(slot * 2 + 0) * kThreads];
    uint4 ue = make_uint4(0u, 0u, 0u, 0u);
    if (accumulate) ue = my[(slot * 2 + 1) * kThreads];
    if (pl < npix) {
      cp_async16(my + (slot * 2 + 0) * kThreads, s + pl * lds + g * 8);
      if (accumulate) cp_async16(my + (slot * 2 + 1) * kThreads, d + pl * ldd + g * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pl += stride;
    if (accumulate) {
      float v[8], e[8];
      unpack8h(us, v);
      unpack8h(ue, e);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += e[k];
      store8<__half>(d + p * ldd + g * 8, v);
    } else {
      *reinterpret_cast<uint4*>(d + p * ldd + g * 8) = us;
    }
    if (++slot == kD) slot = 0;
  }
}

// ------------------------------------------------------------------------------------------
// output head: 1x1 conv (Cin -> 1) + sigmoid, BCE+Dice sums, and their backward
// ------------------------------------------------------------------------------------------
template <typename T, int CIN>
__global__ void __launch_bounds__(kThreads) head_fwd_kernel(const T* __restrict__ x, int ldx,
                                                            const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            float* __restrict__ prob, long long npix) {
  B2U_PDL_PROLOGUE();
  float wr[CIN];
#pragma unroll
  for (int k = 0; k < CIN; ++k) wr[k] = __ldg(w + k);
  const float b = __ldg(bias);
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
       p += (long long)gridDim.x * blockDim.x) {
    float acc = b;
#pragma unroll
    for (int g = 0; g < CIN / 8; ++g) {
      float v[8];
      load8<T>(x + p * ldx + g * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(v[k], wr[g * 8 + k], acc);
    }
    prob[p] = 1.f / (1.f + expf(-acc));
  }
}

__device__ __forceinline__ float bce_term(float t, float p) {
  const float lo = 1e-7f, hi = 1.f - 1e-7f;    // keras.backend.epsilon() clip
  float ph = fminf(fmaxf(p, lo), hi);
  return -(t * logf(ph) + (1.f - t) * log1pf(-ph));
}

__global__ void __launch_bounds__(kThreads) bce_dice_sums_kernel(const float* __restrict__ prob,
                                                                 const float* __restrict__ tgt, long long count,
                                                                 double* __restrict__ sums) {
  B2U_PDL_PROLOGUE();
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    float p = prob[i], t = tgt[i];
    s0 += t * p; s1 += t; s2 += p; s3 += bce_term(t, p);
  }
  __shared__ double sh[4];
  if (threadIdx.x < 4) sh[threadIdx.x] = 0.0;
  __syncthreads();
  double d0 = warp_sum_d((double)s0), d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2),
         d3 = warp_sum_d((double)s3);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sh[0], d0); atomicAdd(&sh[1], d1); atomicAdd(&sh[2], d2); atomicAdd(&sh[3], d3);
  }
  __syncthreads();
  if (threadIdx.x < 4) atomicAdd(&sums[threadIdx.x], sh[threadIdx.x]);
}

__global__ void bce_dice_finalize_kernel(const double* __restrict__ sums, long long count, float* __restrict__ out) {
  B2U_PDL_PROLOGUE();
  double I = sums[0], S = sums[1] + sums[2];
  double dice = (2.0 * I + 1.0) / (S + 1.0);
  out[0] = (float)(0.5 * sums[3] / (double)count + 0.5 * (1.0 - dice));
  out[1] = (float)dice;
}

// Thread = (pixel lane, 8-channel group): 16-byte coalesced loads / stores, the per-pixel loss derivative is
// recomputed by the C/8 threads of a pixel (prob / target are broadcast loads).  Besides dW (sum dl * x) and db it
// can emit `colsum` = per-channel sums of the dx values written (the bias gradient of the conv feeding the head).
template <typename T>
__global__ void __launch_bounds__(kThreads) head_bwd_kernel(
    const float* __restrict__ prob, const float* __restrict__ tgt, const double* __restrict__ sums,
    long long count, const b2u_step_state* __restrict__ st, const T* __restrict__ x, int ldx,
    const float* __restrict__ w, T* __restrict__ dx, int lddx, int x_act, float* __restrict__ dw,
    float* __restrict__ db, long long npix, int C, float* __restrict__ colsum) {
  B2U_PDL_PROLOGUE();
  extern __shared__ float sacc[];          // [2*C + 1]: dW partials, colsum partials, db partial
  for (int i = threadIdx.x; i < 2 * C + 1; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  float wr[8], acc[8], cs[8];
  {
    const int g0 = (threadIdx.x % (C >> 3)) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { wr[k] = __ldg(w + g0 + k); acc[k] = 0.f; cs[k] = 0.f; }
  }
  float accb = 0.f;
  const double I = sums[0], S = sums[1] + sums[2];
  const float inv_s1 = (float)(1.0 / (S + 1.0));
  const float two_i1 = (float)(2.0 * I + 1.0);
  const float scale = st->loss_scale;
  const float half_inv_n = 0.5f / (float)count;
  // two pixels per trip: all loads are issued before the arithmetic (four pixels per trip measured slower: 119 registers,
  // two blocks per SM: 0.099 ms against 0.092 ms).  ncu, round 2: memory-latency bound (5.5 long-scoreboard stalls per
  // issue, 32 % of the DRAM rate), 40 of 183 instructions per pixel were two IEEE divisions -- which cancel:
  //   d(0.5 mean BCE)/dz = 0.5/N (p - t)  inside the clip range   [(-t/p + (1-t)/(1-p)) p (1-p) = p - t]
  //   d(0.5 (1 - dice))/dz = -0.5 (2 t (S+1) - (2I+1)) / (S+1)^2 * p (1 - p)
  constexpr int kPx = 2;
  const int cg = C >> 3;
  const int lanes = kThreads / cg;
  const int g = threadIdx.x % cg, lane_ = threadIdx.x / cg;
  if (lane_ < lanes) {
    for (long long p0 = (long long)blockIdx.x * lanes * kPx + lane_; p0 < npix; p0 += (long long)gridDim.x * lanes * kPx) {
      float v[kPx][8], pr[kPx], tt[kPx];
#pragma unroll
      for (int u = 0; u < kPx; ++u) {
        const long long pu = p0 + (long long)u * lanes;
        const bool on = pu < npix;
        pr[u] = on ? prob[pu] : 0.5f;
        tt[u] = on ? tgt[pu] : 0.f;
        if (on) load8<T>(x + pu * ldx + g * 8, v[u]);
      }
#pragma unroll
      for (int u = 0; u < kPx; ++u) {
        const long long pu = p0 + (long long)u * lanes;
        if (pu >= npix) break;
        const float prv = pr[u], t = tt[u];
        float dl = (prv >= 1e-7f && prv <= 1.f - 1e-7f) ? half_inv_n * (prv - t) : 0.f;
        dl -= 0.5f * (2.f * t - two_i1 * inv_s1) * inv_s1 * (prv * (1.f - prv));
        dl *= scale;
        if (g == 0) accb += dl;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc[k] = fmaf(dl, v[u][k], acc[k]);
          o[k] = dl * wr[k] * act_bwd_from_y(v[u][k], x_act);
          cs[k] += o[k];
        }
        store8<T>(dx + pu * lddx + g * 8, o);
      }
    }
  }
  {
    if (threadIdx.x / (C >> 3) < kThreads / (C >> 3)) {
      group_add8(sacc, acc, C >> 3, threadIdx.x % (C >> 3));
      group_add8(sacc + C, cs, C >> 3, threadIdx.x % (C >> 3));
    }
    // db: only the g == 0 thread of a pixel accumulated it (others hold 0): plain warp sum
    const float sb = warp_sum(accb);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[2 * C], sb);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dw[i], sacc[i]);
    if (colsum != nullptr) atomicAdd(&colsum[i], sacc[C + i]);
  }
  if (threadIdx.x == 0) atomicAdd(db, sacc[2 * C]);
}

// ------------------------------------------------------------------------------------------
// Adam over the flat parameter buffer; step-state bookkeeping; batch gather
// ------------------------------------------------------------------------------------------
// One pass over the (all-reduced) gradient buffer before the update: ANY non-finite value (fp16 overflow of an activation
// gradient under the static loss scale) marks the whole step as skipped -- Adam then leaves parameters and moments
// alone, the step / beta powers do not advance and the loss scale is halved (state_advance_kernel), as dynamic loss
// scaling does; a per-element skip would apply the finite part of a corrupted gradient.  31 MB: ~7 us.
__global__ void __launch_bounds__(kThreads) grad_finite_check_kernel(const float* __restrict__ g, long long n,
                                                                     b2u_step_state* __restrict__ st) {
  B2U_PDL_PROLOGUE();
  bool bad = false;
  const long long nq = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
    const float4 a = reinterpret_cast<const float4*>(g)[q];
    bad |= !(isfinite(a.x) && isfinite(a.y) && isfinite(a.z) && isfinite(a.w));
  }
  for (long long i = (nq << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) bad |= !isfinite(g[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(&st->skip_step, 1u);
}

__global__ void __launch_bounds__(kThreads) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        b2u_step_state* __restrict__ st) {
  B2U_PDL_PROLOGUE();
  if (*reinterpret_cast<volatile const uint32_t*>(&st->skip_step) != 0u) return;     // whole step skipped (set by the check pass)
  // 16-byte accesses (the flat buffers are 16-byte aligned and every tensor is padded to 4 elements), two
  // quads per trip so that eight independent loads are in flight per thread
  const float b1 = st->beta1, b2 = st->beta2, eps = st->eps;
  const float lr_t = st->lr * sqrtf(1.f - st->beta2_pow) / (1.f - st->beta1_pow);
  const float gs = 1.f / (st->loss_scale * st->grad_div);
  bool bad = false;
  const long long nq = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto upd = [&](float gi, float& mi, float& vi, float& pi) {
    gi *= gs;
    if (!isfinite(gi)) { bad = true; return; }
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    pi -= lr_t * mi / (sqrtf(vi) + eps);
  };
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += 2 * stride) {
    const long long q2 = q + stride;
    const bool two = q2 < nq;
    float4 g4[2], m4[2], v4[2], p4[2];
    g4[0] = reinterpret_cast<const float4*>(g)[q]; m4[0] = reinterpret_cast<const float4*>(m)[q];
    v4[0] = reinterpret_cast<const float4*>(v)[q]; p4[0] = reinterpret_cast<const float4*>(p)[q];
    if (two) {
      g4[1] = reinterpret_cast<const float4*>(g)[q2]; m4[1] = reinterpret_cast<const float4*>(m)[q2];
      v4[1] = reinterpret_cast<const float4*>(v)[q2]; p4[1] = reinterpret_cast<const float4*>(p)[q2];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      upd(g4[u].x, m4[u].x, v4[u].x, p4[u].x);
      upd(g4[u].y, m4[u].y, v4[u].y, p4[u].y);
      upd(g4[u].z, m4[u].z, v4[u].z, p4[u].z);
      upd(g4[u].w, m4[u].w, v4[u].w, p4[u].w);
      const long long qq = u ? q2 : q;
      reinterpret_cast<float4*>(m)[qq] = m4[u];
      reinterpret_cast<float4*>(v)[qq] = v4[u];
      reinterpret_cast<float4*>(p)[qq] = p4[u];
    }
  }
  // tail (n is a multiple of 4 for the engine's buffers; kept for direct callers)
  for (long long i = (nq << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float mi = m[i], vi = v[i], pi = p[i];
    upd(g[i], mi, vi, pi);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
  if (bad) atomicOr(&st->overflow, 1u);
}
__global__ void state_advance_kernel(b2u_step_state* st) {
  B2U_PDL_PROLOGUE();
  if (st->skip_step != 0u) {              // skipped step: nothing advances, the loss scale backs off, the host is told
    st->skip_step = 0u;
    st->overflow = 1u;
    st->loss_scale = fmaxf(st->loss_scale * 0.5f, 1.f);
    return;
  }
  st->step += 1;
  st->beta1_pow *= st->beta1;
  st->beta2_pow *= st->beta2;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) gather_batch_kernel(const float* __restrict__ src,
                                                                const int* __restrict__ idx, T* __restrict__ dst,
                                                                long long per_sample, int nb) {
  B2U_PDL_PROLOGUE();
  const long long total = (long long)nb * per_sample;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long b = i / per_sample;
    long long e = i - b * per_sample;
    long long s = idx ? (long long)idx[b] : b;
    stf<T>(dst + i, src[s * per_sample + e]);
  }
}

// the same gather with the channel count padded from c to cpad (zeros): an inference plan whose first conv has
// 2..15 input channels (Task-2's 224 x 224 x 3 slices) pads them to 16 so that the tcgen05 kernel takes the layer
template <typename T>
__global__ void __launch_bounds__(kThreads) gather_batch_pad_kernel(const float* __restrict__ src,
                                                                    const int* __restrict__ idx, T* __restrict__ dst,
                                                                    long long pix_per_sample, int c, int cpad, int nb) {
  B2U_PDL_PROLOGUE();
  // thread = pixel: c contiguous floats in, cpad / 8 sixteen-byte (fp16) stores out (cpad % 8 == 0, c <= 8 handled in
  // the first vector, the rest of the padding is zeros)
  const long long total = (long long)nb * pix_per_sample;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long b = pix / pix_per_sample, e = pix - b * pix_per_sample;
    const long long s = idx ? (long long)idx[b] : b;
    const float* sp = src + (s * pix_per_sample + e) * c;
    T* dp = dst + pix * cpad;
    for (int v0 = 0; v0 < cpad; v0 += 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (v0 + k < c) ? sp[v0 + k] : 0.f;
      store8<T>(dp + v0, v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// sm.metrics threshold sweep: one pass over (p, t) for all thresholds
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) threshold_counts_kernel(const float* __restrict__ prob,
                                                                    const float* __restrict__ tgt, long long count,
                                                                    const float* __restrict__ thr, int nthr,
                                                                    double* __restrict__ tp, double* __restrict__ spr,
                                                                    double* __restrict__ sgt) {
  B2U_PDL_PROLOGUE();
  extern __shared__ float sh[];   // [nthr] tp, [nthr] cnt, [1] gt
  for (int i = threadIdx.x; i < 2 * nthr + 1; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  float gt = 0.f;
  const long long start = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // all lanes of a warp iterate together (count is padded by predication)
  for (long long i0 = start - (threadIdx.x & 31); i0 < count; i0 += stride) {
    long long i = i0 + (threadIdx.x & 31);
    bool ok = i < count;
    float p = ok ? prob[i] : -1.f, t = ok ? tgt[i] : 0.f;
    gt += t;
    for (int k = 0; k < nthr; ++k) {
      bool pr = ok && (p > __ldg(thr + k));
      unsigned b = __ballot_sync(0xffffffffu, pr);
      if (b == 0u) continue;
      float s = warp_sum(pr ? t : 0.f);
      if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[k], s); atomicAdd(&sh[nthr + k], (float)__popc(b)); }
    }
  }
  gt = warp_sum(gt);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sh[2 * nthr], gt);
  __syncthreads();
  for (int k = threadIdx.x; k < nthr; k += blockDim.x) {
    atomicAdd(&tp[k], (double)sh[k]);
    atomicAdd(&spr[k], (double)sh[nthr + k]);
  }
  if (threadIdx.x == 0) atomicAdd(sgt, (double)sh[2 * nthr]);
}

// ------------------------------------------------------------------------------------------
// Dense + classifier loss (Task-2 head; small, HBM-bound)
// ------------------------------------------------------------------------------------------
template <typename T, int M>
__global__ void __launch_bounds__(kThreads) dense_fwd_kernel(const T* __restrict__ x, int K,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ bias, int act,
                                                             float* __restrict__ y) {
  B2U_PDL_PROLOGUE();
  // one block per sample; threads split K; each keeps M partial outputs (y is always fp32)
  __shared__ float sacc[M];
  for (int i = threadIdx.x; i < M; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.x;
  float acc[M];
#pragma unroll
  for (int j = 0; j < M; ++j) acc[j] = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float xv = ldf<T>(x + (long long)n * K + k);
#pragma unroll
    for (int j = 0; j < M; ++j) acc[j] = fmaf(xv, __ldg(w + (long long)k * M + j), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < M; ++j) {
    float s = warp_sum(acc[j]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[j], s);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    float v = sacc[j] + bias[j];
    if (act == B2U_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    else v = act_fwd(v, act);
    y[(long long)n * M + j] = v;
  }
}

// Dense(K -> 32) for large K (the classifier's Flatten -> Dense(32): K = 50176, T2:776): split-K over the grid.  A block
// owns kDK consecutive k for a group of up to kDN samples: its slice of the kernel (kDK x 32 fp32) and of the activations
// (kDN x kDK) is staged in shared memory once, thread (sample group of 4, unit m) keeps 4 partial sums, and the block adds
// its kDN x 32 partial outputs to an fp32 accumulator with atomics; dense_finalize_kernel applies bias + activation.
// (The one-block-per-sample kernel above re-reads the whole 6.4 MB kernel per sample with one scalar load per weight:
// 0.84 ms for batch 64 on B200 against ~10 us for this one.)
constexpr int kDK = 128, kDN = 32;       // 16 KB of kernel + 16.5 KB of activations per block
template <typename T>
__global__ void __launch_bounds__(kThreads) dense32_splitk_kernel(const T* __restrict__ x, int K,
                                                                  const float* __restrict__ w, int N,
                                                                  float* __restrict__ acc) {
  B2U_PDL_PROLOGUE();
  __shared__ float ws_[kDK][32];
  __shared__ float xs[kDN][kDK + 1];
  const int k0 = blockIdx.x * kDK, n0 = blockIdx.y * kDN;
  const int kc = min(kDK, K - k0), nc = min(kDN, N - n0);
  for (int i = threadIdx.x; i < kDK * 32; i += kThreads) {
    const int k = i >> 5;
    ws_[k][i & 31] = k < kc ? __ldg(w + (long long)(k0 + k) * 32 + (i & 31)) : 0.f;
  }
  for (int i = threadIdx.x; i < kDN * kDK; i += kThreads) {
    const int n = i / kDK, k = i % kDK;
    xs[n][k] = (n < nc && k < kc) ? ldf<T>(x + (long long)(n0 + n) * K + k0 + k) : 0.f;
  }
  __syncthreads();
  constexpr int SPW = kDN / (kThreads / 32);                       // samples per warp
  const int m = threadIdx.x & 31, ng = threadIdx.x >> 5;
  float a[SPW];
#pragma unroll
  for (int q = 0; q < SPW; ++q) a[q] = 0.f;
#pragma unroll 4
  for (int k = 0; k < kDK; ++k) {
    const float wv = ws_[k][m];                                    // conflict-free; xs reads are warp broadcasts
#pragma unroll
    for (int q = 0; q < SPW; ++q) a[q] = fmaf(xs[ng * SPW + q][k], wv, a[q]);
  }
#pragma unroll
  for (int q = 0; q < SPW; ++q)
    if (ng * SPW + q < nc) atomicAdd(acc + (long long)(n0 + ng * SPW + q) * 32 + m, a[q]);
}

__global__ void __launch_bounds__(kThreads) dense_finalize_kernel(const float* __restrict__ acc,
                                                                  const float* __restrict__ bias, int act, int M,
                                                                  long long total, float* __restrict__ y) {
  B2U_PDL_PROLOGUE();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = acc[i] + bias[i % M];
  if (act == B2U_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
  else v = act_fwd(v, act);
  y[i] = v;
}

// dpre[n][j] = dy[n][j] * act'(y[n][j])   (sigmoid handled by the loss kernel: act NONE there)
template <typename T, int M>
__global__ void __launch_bounds__(kThreads) dense_bwd_kernel(const T* __restrict__ x, int K,
                                                             const float* __restrict__ w, const float* __restrict__ y,
                                                             int act, const float* __restrict__ dy, T* __restrict__ dx,
                                                             const T* __restrict__ mask, int mask_act,
                                                             float* __restrict__ dw, float* __restrict__ db, int N) {
  B2U_PDL_PROLOGUE();
  // grid over K chunks: each thread owns one k and loops over samples (N small) -> dw row + dx column
  extern __shared__ float dpre[];   // [N][M]
  for (int i = threadIdx.x; i < N * M; i += blockDim.x) {
    dpre[i] = dy[i] * act_bwd_from_y(y[i], act);
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += dpre[n * M + j];
      db[j] += s;
    }
  }
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float wr[M], gw[M];
#pragma unroll
  for (int j = 0; j < M; ++j) { wr[j] = w[(long long)k * M + j]; gw[j] = 0.f; }
  for (int n = 0; n < N; ++n) {
    float xv = ldf<T>(x + (long long)n * K + k);
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      float dp = dpre[n * M + j];
      gw[j] = fmaf(xv, dp, gw[j]);
      d = fmaf(dp, wr[j], d);
    }
    if (dx != nullptr) {
      if (mask != nullptr) d *= act_bwd_from_y(ldf<T>(mask + (long long)n * K + k), mask_act);
      stf<T>(dx + (long long)n * K + k, d);
    }
  }
#pragma unroll
  for (int j = 0; j < M; ++j) dw[(long long)k * M + j] += gw[j];
}

__global__ void bce_fwd_kernel(const float* __restrict__ prob, const float* __restrict__ tgt,
                               const float* __restrict__ sw, int n, float* __restrict__ out) {
  B2U_PDL_PROLOGUE();
  // single block; keras: mean over batch of w_i * bce_i
  __shared__ float sh;
  if (threadIdx.x == 0) sh = 0.f;
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (sw ? sw[i] : 1.f) * bce_term(tgt[i], prob[i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sh, s);
  __syncthreads();
  if (threadIdx.x == 0) out[0] = sh / (float)n;
}

template <typename T>
__global__ void bce_sigmoid_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ tgt,
                                       const float* __restrict__ sw, int n, const b2u_step_state* __restrict__ st,
                                       T* __restrict__ dlogit) {
  B2U_PDL_PROLOGUE();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p = prob[i], t = tgt[i], g = 0.f;
  if (p >= 1e-7f && p <= 1.f - 1e-7f) g = (-t / p + (1.f - t) / (1.f - p)) * p * (1.f - p);
  stf<T>(dlogit + i, g * (sw ? sw[i] : 1.f) / (float)n * st->loss_scale);
}

}  // namespace

// ==========================================================================================
// C-ABI wrappers
// ==========================================================================================
#define DISPATCH_T(dt, ...)                                              \
  if ((dt) == B2U_F32) { using T = float; __VA_ARGS__; }                 \
  else if ((dt) == B2U_F16) { using T = __half; __VA_ARGS__; }           \
  else { b2u_set_error("bad dtype %d", (int)(dt)); return B2U_ERR_ARG; }

#define REQ_VEC8(c, ...)                                                                           \
  B2U_REQUIRE((c) > 0 && (c) % 8 == 0, "%s: channel count %d must be a positive multiple of 8", __func__, (int)(c))

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int g_b2u_bn_async = 1;     // 1: cp.async versions of the fp16 BatchNorm apply / backward apply kernels

extern "C" int b2u_state_advance(b2u_step_state* d_state, void* stream) {
  B2U_LAUNCH(state_advance_kernel, 1, 1, 0, stream, d_state);
  return B2U_OK;
}

extern "C" int b2u_bn_stats(int dt, const void* x, int ldx, int c, long long npix, double* sums, void* stream) {
  return b2u_bn_stats_off(dt, x, ldx, c, npix, sums, c, stream, 0.f, 0, nullptr, nullptr);
}

// statistics of a tensor into a (possibly wider) BN sums buffer: sums[i], sums[sq_off + i]
int b2u_bn_stats_off(int dt, const void* x, int ldx, int c, long long npix, double* sums, int sq_off, void* stream,
                     float p_drop, int op_id, const b2u_step_state* d_state, void* drop_bits) {
  REQ_VEC8(c);
  B2U_REQUIRE(c <= 2048 && ldx % 8 == 0 && aligned16(x), "bn_stats: c<=2048, ld%%8==0, 16B-aligned base required");
  int lanes = kThreads / (c / 8);
  int grid = stream_grid((npix + 1) / 2, lanes, 4);
  size_t smem = 2 * (size_t)c * sizeof(float);
  if (p_drop > 0.f) {
    B2U_REQUIRE(d_state != nullptr && drop_bits != nullptr && p_drop < 1.f,
                "bn_stats: dropout needs a step state, a keep-mask buffer and 0 <= p < 1");
    DISPATCH_T(dt, B2U_LAUNCH((bn_reduce_kernel<T, false, true>), grid, kThreads, smem, stream, (const T*)x, ldx,
                              (const T*)nullptr, 0, c, npix, (const float*)nullptr, (const float*)nullptr, sums, sq_off,
                              d_state, p_drop, op_id, (uint8_t*)drop_bits));
  } else {
    DISPATCH_T(dt, B2U_LAUNCH((bn_reduce_kernel<T, false, false>), grid, kThreads, smem, stream, (const T*)x, ldx,
                              (const T*)nullptr, 0, c, npix, (const float*)nullptr, (const float*)nullptr, sums, sq_off,
                              (const b2u_step_state*)nullptr, 0.f, 0, (uint8_t*)nullptr));
  }
  return B2U_OK;
}

extern "C" int b2u_bn_finalize(const double* sums, long long count, const float* gamma, const float* beta,
                               float* moving_mean, float* moving_var, float momentum, float eps, int training,
                               float* scale, float* shift, float* save_mean, float* save_invstd, int c,
                               void* stream) {
  B2U_REQUIRE(c > 0, "bn_finalize: c must be > 0");
  B2U_LAUNCH(bn_finalize_kernel, b2u_cdiv(c, 128), 128, 0, stream, sums, count, gamma, beta, moving_mean,
             moving_var, momentum, eps, training, scale, shift, save_mean, save_invstd, c);
  return B2U_OK;
}

extern "C" int b2u_bn_apply(int dt, const void* x, int ldx, void* y, int ldy, int c, long long npix,
                            const float* scale, const float* shift, double* out_stats, int out_sq_off,
                            void* stream) {
  return b2u_bn_apply_split(dt, x, ldx, nullptr, 0, c, y, ldy, c, npix, scale, shift, out_stats, out_sq_off, stream, 0.f,
                            nullptr);
}

// channels [0, split) from x, [split, c) from x2 (op lists only: a two-input concatenate kept as two dense tensors)
int b2u_bn_apply_split(int dt, const void* x, int ldx, const void* x2, int ldx2, int split, void* y, int ldy, int c,
                       long long npix, const float* scale, const float* shift, double* out_stats, int out_sq_off,
                       void* stream, float p_drop, const void* drop_bits) {
  REQ_VEC8(c);
  B2U_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y), "bn_apply: alignment");
  B2U_REQUIRE(c <= 2048, "bn_apply: c <= 2048");
  if (x2 == nullptr) split = c;
  B2U_REQUIRE(split == c || (split > 0 && split < c && split % 8 == 0 && ldx2 % 8 == 0 && aligned16(x2)),
              "bn_apply: bad second source (split=%d of %d)", split, c);
  B2U_REQUIRE(p_drop == 0.f || (drop_bits != nullptr && p_drop > 0.f && p_drop < 1.f && split == c),
              "bn_apply: dropout needs the keep mask, 0 <= p < 1 and a single source");
  int grid = lane_grid(npix, c, out_stats != nullptr ? 4 : 8);
  size_t smem = out_stats != nullptr ? 2 * (size_t)c * sizeof(float) : 0;
  if (dt == B2U_F16 && g_b2u_bn_async && out_stats == nullptr) {                           // the cp.async version
    B2U_LAUNCH(bn_apply_async_kernel, grid, kThreads, (size_t)kAA * kThreads * sizeof(uint4), stream, (const __half*)x, ldx,
               (__half*)y, ldy, c, npix, scale, shift, (const __half*)x2, ldx2, split, (const uint8_t*)drop_bits, p_drop);
    return B2U_OK;
  }
  DISPATCH_T(dt, B2U_LAUNCH(bn_apply_kernel<T>, grid, kThreads, smem, stream, (const T*)x, ldx, (T*)y, ldy, c, npix,
                            scale, shift, out_stats, out_sq_off, (const T*)x2, ldx2, split, (const uint8_t*)drop_bits,
                            p_drop));
  return B2U_OK;
}

int b2u_relu_bits(int dt, const void* y, int ldy, int c, long long npix, void* bits, void* stream) {
  REQ_VEC8(c);
  B2U_REQUIRE(ldy % 8 == 0 && aligned16(y) && c <= 2048 && bits != nullptr, "relu_bits: args");
  DISPATCH_T(dt, B2U_LAUNCH(relu_bits_kernel<T>, lane_grid(npix, c), kThreads, 0, stream, (const T*)y, ldy, c, npix,
                            (uint8_t*)bits));
  return B2U_OK;
}

int b2u_apply_relu_bits(int dt, void* dx, int lddx, int c, long long npix, const void* bits, void* stream) {
  REQ_VEC8(c);
  B2U_REQUIRE(lddx % 8 == 0 && aligned16(dx) && c <= 2048 && bits != nullptr, "apply_relu_bits: args");
  DISPATCH_T(dt, B2U_LAUNCH(apply_bits_kernel<T>, lane_grid(npix, c), kThreads, 0, stream, (T*)dx, lddx, c, npix,
                            (const uint8_t*)bits));
  return B2U_OK;
}

extern "C" int b2u_bn_apply_pool(int dt, const void* x, int ldx, void* y, int ldy, int c, int n, int h, int wd,
                                 const float* scale, const float* shift, double* out_stats, int out_sq_off, void* yp,
                                 int ldp, float p_drop, int op_id, const b2u_step_state* d_state, void* stream) {
  REQ_VEC8(c);
  B2U_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldp % 8 == 0 && aligned16(x) && aligned16(y) && aligned16(yp),
              "bn_apply_pool: alignment");
  B2U_REQUIRE(h % 2 == 0 && wd % 2 == 0 && c <= 2048 && (long long)n * h * wd < (1LL << 31), "bn_apply_pool: shape");
  B2U_REQUIRE(p_drop == 0.f || d_state != nullptr, "bn_apply_pool: dropout needs a step state");
  B2U_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "bn_apply_pool: bad dropout rate");
  int grid = lane_grid((long long)n * (h / 2) * (wd / 2), c, out_stats != nullptr ? 4 : 8);
  size_t smem = out_stats != nullptr ? 2 * (size_t)c * sizeof(float) : 0;
  if (dt == B2U_F16 && g_b2u_bn_async) {                    // the cp.async version (64 KB ring + statistics partials)
    static bool attr = false;
    if (!attr) {
      B2U_CHECK_CUDA(cudaFuncSetAttribute(bn_apply_pool_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr = true;
    }
    B2U_LAUNCH(bn_apply_pool_async_kernel, grid, kThreads, (size_t)kPD * 4 * kThreads * sizeof(uint4) + 2 * (size_t)c * sizeof(float),
               stream, (const __half*)x, ldx, (__half*)y, ldy, c, n, h, wd, scale, shift, out_stats, out_sq_off, (__half*)yp, ldp,
               p_drop, op_id, d_state);
    return B2U_OK;
  }
  DISPATCH_T(dt, B2U_LAUNCH(bn_apply_pool_kernel<T>, grid, kThreads, smem, stream, (const T*)x, ldx, (T*)y, ldy, c, n, h,
                            wd, scale, shift, out_stats, out_sq_off, (T*)yp, ldp, p_drop, op_id, d_state));
  return B2U_OK;
}

extern "C" int b2u_bn_bwd_sums_from_wgrad(const float* w, const float* dw, const float* colsum, const float* gamma,
                                          const float* beta, double* sums, int c, int cout, int taps, void* stream) {
  B2U_REQUIRE(w && dw && colsum && gamma && beta && sums && c > 0 && cout > 0 && taps > 0, "bn_bwd_sums_from_wgrad: args");
  B2U_LAUNCH(bn_bwd_sums_wgrad_kernel, c, 128, 0, stream, w, dw, colsum, gamma, beta, sums, c, cout, taps);
  return B2U_OK;
}

extern "C" int b2u_bn_bwd_reduce(int dt, const void* dy, int lddy, const void* x, int ldx, int c, long long npix,
                                 const float* save_mean, const float* save_invstd, double* sums, void* stream) {
  return b2u_bn_bwd_reduce_off(dt, dy, lddy, x, ldx, c, npix, save_mean, save_invstd, sums, c, stream, 0.f, nullptr);
}

// the same into a (possibly wider) sums buffer: sums[i], sums[sq_off + i] (one half of a split concatenate)
int b2u_bn_bwd_reduce_off(int dt, const void* dy, int lddy, const void* x, int ldx, int c, long long npix,
                          const float* save_mean, const float* save_invstd, double* sums, int sq_off, void* stream,
                          float p_drop, const void* drop_bits) {
  REQ_VEC8(c);
  B2U_REQUIRE(c <= 2048 && ldx % 8 == 0 && lddy % 8 == 0 && aligned16(x) && aligned16(dy), "bn_bwd_reduce: alignment");
  int lanes = kThreads / (c / 8);
  int grid = stream_grid((npix + 1) / 2, lanes, 4);
  size_t smem = 2 * (size_t)c * sizeof(float);
  if (p_drop > 0.f) {
    B2U_REQUIRE(drop_bits != nullptr && p_drop < 1.f, "bn_bwd_reduce: dropout needs the keep mask and 0 <= p < 1");
    DISPATCH_T(dt, B2U_LAUNCH((bn_reduce_kernel<T, true, true>), grid, kThreads, smem, stream, (const T*)dy, lddy,
                              (const T*)x, ldx, c, npix, save_mean, save_invstd, sums, sq_off,
                              (const b2u_step_state*)nullptr, p_drop, 0, (uint8_t*)const_cast<void*>(drop_bits)));
  } else {
    DISPATCH_T(dt, B2U_LAUNCH((bn_reduce_kernel<T, true, false>), grid, kThreads, smem, stream, (const T*)dy, lddy,
                              (const T*)x, ldx, c, npix, save_mean, save_invstd, sums, sq_off,
                              (const b2u_step_state*)nullptr, 0.f, 0, (uint8_t*)nullptr));
  }
  return B2U_OK;
}

extern "C" int b2u_bn_bwd_apply(int dt, const void* dy, int lddy, const void* x, int ldx, void* dx, int lddx, int c,
                                long long npix, long long count, const float* gamma, const float* save_mean,
                                const float* save_invstd, const double* sums, float* dgamma, float* dbeta,
                                const void* mask, int ldmask, int mask_act, void* stream) {
  return b2u_bn_bwd_apply_cs(dt, dy, lddy, x, ldx, dx, lddx, c, npix, count, gamma, save_mean, save_invstd, sums, dgamma,
                             dbeta, mask, ldmask, mask_act, nullptr, stream);
}

int b2u_bn_bwd_apply_cs(int dt, const void* dy, int lddy, const void* x, int ldx, void* dx, int lddx, int c,
                        long long npix, long long count, const float* gamma, const float* save_mean,
                        const float* save_invstd, const double* sums, float* dgamma, float* dbeta, const void* mask,
                        int ldmask, int mask_act, float* colsum, void* stream, const void* x2, int ldx2, void* dx2,
                        int lddx2, int split, float p_drop, const void* drop_bits) {
  REQ_VEC8(c);
  B2U_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && (mask == nullptr || ldmask % 8 == 0),
              "bn_bwd_apply: alignment");
  B2U_REQUIRE(c <= 2048, "bn_bwd_apply: c <= 2048");
  if (x2 == nullptr) split = c;
  B2U_REQUIRE(split == c || (split > 0 && split < c && split % 8 == 0 && ldx2 % 8 == 0 && lddx2 % 8 == 0 && dx2 != nullptr &&
                             aligned16(x2) && aligned16(dx2)),
              "bn_bwd_apply: bad second tensor (split=%d of %d)", split, c);
  B2U_REQUIRE(p_drop == 0.f || (drop_bits != nullptr && p_drop > 0.f && p_drop < 1.f && split == c),
              "bn_bwd_apply: dropout needs the keep mask, 0 <= p < 1 and a single tensor");
  int grid = lane_grid(npix, c);
  // fp16, one tensor, no dropout, activation mask (if any) = the BN input itself: the cp.async version
  if (dt == B2U_F16 && g_b2u_bn_async && aligned16(dy) && aligned16(x) && aligned16(dx) &&
      (mask == nullptr || (mask == x && ldmask == ldx && split == c))) {
    const size_t ring_bytes = (size_t)kAD * 2 * kThreads * sizeof(uint4);
    const int ma = mask != nullptr ? mask_act : B2U_ACT_NONE;
    if (colsum != nullptr) {
      B2U_LAUNCH(bn_bwd_apply_async_kernel<true>, grid, kThreads, ring_bytes + c * sizeof(float), stream, (const __half*)dy,
                 lddy, (const __half*)x, ldx, (__half*)dx, lddx, c, npix, count, gamma, save_mean, save_invstd, sums, dgamma,
                 dbeta, ma, colsum, (const __half*)x2, ldx2, (__half*)dx2, lddx2, split, (const uint8_t*)drop_bits, p_drop);
    } else {
      B2U_LAUNCH(bn_bwd_apply_async_kernel<false>, grid, kThreads, ring_bytes, stream, (const __half*)dy, lddy,
                 (const __half*)x, ldx, (__half*)dx, lddx, c, npix, count, gamma, save_mean, save_invstd, sums, dgamma, dbeta, ma,
                 colsum, (const __half*)x2, ldx2, (__half*)dx2, lddx2, split, (const uint8_t*)drop_bits, p_drop);
    }
    return B2U_OK;
  }
  if (colsum != nullptr) {
    DISPATCH_T(dt, B2U_LAUNCH((bn_bwd_apply_kernel<T, true, true>), grid, kThreads, c * sizeof(float), stream, (const T*)dy,
                              lddy, (const T*)x, ldx, (T*)dx, lddx, c, npix, count, gamma, save_mean, save_invstd, sums,
                              dgamma, dbeta, (const T*)mask, ldmask, mask_act, colsum, (const T*)x2, ldx2, (T*)dx2, lddx2,
                              split, (const uint8_t*)drop_bits, p_drop));
  } else if (p_drop > 0.f) {
    DISPATCH_T(dt, B2U_LAUNCH((bn_bwd_apply_kernel<T, false, true>), grid, kThreads, 0, stream, (const T*)dy,
                              lddy, (const T*)x, ldx, (T*)dx, lddx, c, npix, count, gamma, save_mean, save_invstd, sums,
                              dgamma, dbeta, (const T*)mask, ldmask, mask_act, colsum, (const T*)x2, ldx2, (T*)dx2, lddx2,
                              split, (const uint8_t*)drop_bits, p_drop));
  } else {
    DISPATCH_T(dt, B2U_LAUNCH((bn_bwd_apply_kernel<T, false, false>), grid, kThreads, 0, stream, (const T*)dy,
                              lddy, (const T*)x, ldx, (T*)dx, lddx, c, npix, count, gamma, save_mean, save_invstd, sums,
                              dgamma, dbeta, (const T*)mask, ldmask, mask_act, colsum, (const T*)x2, ldx2, (T*)dx2, lddx2,
                              split, (const uint8_t*)drop_bits, p_drop));
  }
  return B2U_OK;
}

extern "C" int b2u_maxpool_fwd(int dt, const void* x, int ldx, void* y, int ldy, int c, int n, int h, int wd,
                               float p_drop, int op_id, const b2u_step_state* d_state, void* stream) {
  REQ_VEC8(c);
  B2U_REQUIRE(h % 2 == 0 && wd % 2 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "maxpool_fwd: even H,W and ld%%8==0");
  B2U_REQUIRE(p_drop == 0.f || d_state != nullptr, "maxpool_fwd: dropout needs a step state");
  B2U_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "maxpool_fwd: bad dropout rate");
  B2U_REQUIRE(c <= 2048 && (long long)n * h * wd < (1LL << 31), "maxpool_fwd: tensor too large");
  int grid = lane_grid((long long)n * (h / 2) * (wd / 2), c);
  DISPATCH_T(dt, B2U_LAUNCH(maxpool_fwd_kernel<T>, grid, kThreads, 0, stream, (const T*)x, ldx, (T*)y, ldy, c, n, h,
                            wd, p_drop, op_id, d_state));
  return B2U_OK;
}

extern "C" int b2u_maxpool_bwd(int dt, const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int c,
                               int n, int h, int wd, float p_drop, int op_id, const b2u_step_state* d_state,
                               int accumulate, double* bn_sums, const float* bn_gamma, const float* bn_beta,
                               void* stream) {
  REQ_VEC8(c);
  B2U_REQUIRE(h % 2 == 0 && wd % 2 == 0 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0, "maxpool_bwd: shape");
  B2U_REQUIRE(p_drop == 0.f || d_state != nullptr, "maxpool_bwd: dropout needs a step state");
  B2U_REQUIRE(c <= 2048 && (long long)n * h * wd < (1LL << 31), "maxpool_bwd: tensor too large");
  int grid = lane_grid((long long)n * (h / 2) * (wd / 2), c);
  B2U_REQUIRE(bn_sums == nullptr || (bn_gamma != nullptr && bn_beta != nullptr), "maxpool_bwd: fused BN statistics need gamma/beta");
  if (bn_sums != nullptr && grid > 4 * B2U_NUM_SMS) grid = 4 * B2U_NUM_SMS;     // fewer block epilogues
  size_t smem = bn_sums != nullptr ? 2 * (size_t)c * sizeof(float) : 0;
  if (dt == B2U_F16 && g_b2u_bn_async && aligned16(x) && aligned16(dy) && aligned16(dx)) {   // the cp.async version (72 KB ring)
    static bool attr = false;
    if (!attr) {
      B2U_CHECK_CUDA(cudaFuncSetAttribute(maxpool_bwd_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr = true;
    }
    B2U_LAUNCH(maxpool_bwd_async_kernel, grid, kThreads, (size_t)kMD * 9 * kThreads * sizeof(uint4) + 2 * (size_t)c * sizeof(float),
               stream, (const __half*)x, ldx, (const __half*)dy, lddy, (__half*)dx, lddx, c, n, h, wd, p_drop, op_id, d_state,
               accumulate, bn_sums, bn_gamma, bn_beta);
    return B2U_OK;
  }
  DISPATCH_T(dt, B2U_LAUNCH(maxpool_bwd_kernel<T>, grid, kThreads, smem, stream, (const T*)x, ldx, (const T*)dy, lddy,
                            (T*)dx, lddx, c, n, h, wd, p_drop, op_id, d_state, accumulate, bn_sums, bn_gamma, bn_beta));
  return B2U_OK;
}

extern "C" int b2u_dropout_fwd(int dt, const void* x, int ldx, void* y, int ldy, int c, long long npix, float p,
                               int op_id, const b2u_step_state* d_state, void* stream) {
  REQ_VEC8(c);
  B2U_REQUIRE(p >= 0.f && p < 1.f && d_state != nullptr && ldx % 8 == 0 && ldy % 8 == 0, "dropout_fwd: args");
  B2U_REQUIRE(c <= 2048, "dropout: c <= 2048");
  int grid = lane_grid(npix, c);
  DISPATCH_T(dt, B2U_LAUNCH(dropout_kernel<T>, grid, kThreads, 0, stream, (const T*)x, ldx, (T*)y, ldy, c, npix, p,
                            op_id, d_state, (const T*)nullptr, 0, 0));
  return B2U_OK;
}

extern "C" int b2u_dropout_bwd(int dt, const void* dy, int lddy, void* dx, int lddx, int c, long long npix, float p,
                               int op_id, const b2u_step_state* d_state, const void* mask, int ldmask, int mask_act,
                               void* stream) {
  REQ_VEC8(c);
  B2U_REQUIRE(p >= 0.f && p < 1.f && d_state != nullptr && lddy % 8 == 0 && lddx % 8 == 0, "dropout_bwd: args");
  B2U_REQUIRE(c <= 2048, "dropout: c <= 2048");
  int grid = lane_grid(npix, c);
  DISPATCH_T(dt, B2U_LAUNCH(dropout_kernel<T>, grid, kThreads, 0, stream, (const T*)dy, lddy, (T*)dx, lddx, c, npix,
                            p, op_id, d_state, (const T*)mask, ldmask, mask_act));
  return B2U_OK;
}

extern "C" int b2u_copy_slice(int dt, const void* src, int ldsrc, void* dst, int lddst, int c, long long npix,
                              int accumulate, void* stream) {
  REQ_VEC8(c);
  B2U_REQUIRE(ldsrc % 8 == 0 && lddst % 8 == 0, "copy_slice: ld%%8==0 required");
  B2U_REQUIRE(c <= 2048, "copy_slice: c <= 2048");
  int grid = lane_grid(npix, c);
  if (dt == B2U_F16 && g_b2u_bn_async && aligned16(src) && aligned16(dst)) {
    B2U_LAUNCH(copy_slice_async_kernel, grid, kThreads, (size_t)kAA * kThreads * sizeof(uint4), stream, (const __half*)src, ldsrc,
               (__half*)dst, lddst, c, npix, accumulate);
    return B2U_OK;
  }
  DISPATCH_T(dt, B2U_LAUNCH(copy_slice_kernel<T>, grid, kThreads, 0, stream, (const T*)src, ldsrc, (T*)dst, lddst, c,
                            npix, accumulate));
  return B2U_OK;
}

extern "C" int b2u_head_fwd(int dt, const void* x, int ldx, int cin, const float* w, const float* bias, float* prob,
                            long long npix, void* stream) {
  B2U_REQUIRE(cin == 32 || cin == 16 || cin == 64, "head_fwd: cin must be 16, 32 or 64 (got %d)", cin);
  B2U_REQUIRE(ldx % 8 == 0, "head_fwd: ld%%8==0 required");
  int grid = stream_grid(npix);
#define HEAD_FWD(CI) DISPATCH_T(dt, B2U_LAUNCH((head_fwd_kernel<T, CI>), grid, kThreads, 0, stream, (const T*)x, ldx, w, bias, prob, npix))
  if (cin == 32) { HEAD_FWD(32); } else if (cin == 16) { HEAD_FWD(16); } else { HEAD_FWD(64); }
#undef HEAD_FWD
  return B2U_OK;
}

extern "C" int b2u_bce_dice_sums(const float* prob, const float* target, long long count, double* sums,
                                 void* stream) {
  int grid = stream_grid(count, kThreads, 4);
  B2U_LAUNCH(bce_dice_sums_kernel, grid, kThreads, 0, stream, prob, target, count, sums);
  return B2U_OK;
}

extern "C" int b2u_bce_dice_finalize(const double* sums, long long count, float* out, void* stream) {
  B2U_LAUNCH(bce_dice_finalize_kernel, 1, 1, 0, stream, sums, count, out);
  return B2U_OK;
}

extern "C" int b2u_head_bwd(int dt, const float* prob, const float* target, const double* sums, long long count,
                            const b2u_step_state* d_state, const void* x, int ldx, int cin, const float* w,
                            void* dx, int lddx, int x_act, float* dw, float* db, long long npix, void* stream) {
  return b2u_head_bwd_cs(dt, prob, target, sums, count, d_state, x, ldx, cin, w, dx, lddx, x_act, dw, db, npix, nullptr,
                         stream);
}

int b2u_head_bwd_cs(int dt, const float* prob, const float* target, const double* sums, long long count,
                    const b2u_step_state* d_state, const void* x, int ldx, int cin, const float* w, void* dx, int lddx,
                    int x_act, float* dw, float* db, long long npix, float* colsum, void* stream) {
  B2U_REQUIRE(cin % 8 == 0 && cin >= 8 && cin <= 256, "head_bwd: cin must be a multiple of 8 in [8, 256] (got %d)", cin);
  B2U_REQUIRE(ldx % 8 == 0 && lddx % 8 == 0 && d_state != nullptr, "head_bwd: args");
  int grid = lane_grid((npix + 1) / 2, cin);
  DISPATCH_T(dt, B2U_LAUNCH(head_bwd_kernel<T>, grid, kThreads, (2 * cin + 1) * sizeof(float), stream, prob, target,
                            sums, count, d_state, (const T*)x, ldx, w, (T*)dx, lddx, x_act, dw, db, npix, cin, colsum));
  return B2U_OK;
}

extern "C" int b2u_adam(float* params, const float* grads, float* m, float* v, long long n, b2u_step_state* d_state,
                        void* stream) {
  B2U_REQUIRE(n >= 0 && d_state != nullptr, "adam: args");
  if (n == 0) return B2U_OK;
  B2U_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam: buffers must be 16-byte aligned");
  B2U_LAUNCH(grad_finite_check_kernel, stream_grid((n + 7) / 8), kThreads, 0, stream, grads, n, d_state);
  B2U_LAUNCH(adam_kernel, stream_grid((n + 7) / 8), kThreads, 0, stream, params, grads, m, v, n, d_state);
  return B2U_OK;
}

extern "C" int b2u_gather_batch(int dt, const float* src, const int* idx, void* dst, long long per_sample, int nb,
                                void* stream) {
  B2U_REQUIRE(nb > 0 && per_sample > 0, "gather_batch: empty batch");
  int grid = stream_grid((long long)nb * per_sample);
  DISPATCH_T(dt, B2U_LAUNCH(gather_batch_kernel<T>, grid, kThreads, 0, stream, src, idx, (T*)dst, per_sample, nb));
  return B2U_OK;
}

extern "C" int b2u_gather_batch_pad(int dt, const float* src, const int* idx, void* dst, long long pix_per_sample, int c,
                                    int cpad, int nb, void* stream) {
  B2U_REQUIRE(nb > 0 && pix_per_sample > 0 && c > 0 && cpad >= c && cpad % 8 == 0 && aligned16(dst), "gather_batch_pad: args");
  int grid = stream_grid((long long)nb * pix_per_sample);
  DISPATCH_T(dt, B2U_LAUNCH(gather_batch_pad_kernel<T>, grid, kThreads, 0, stream, src, idx, (T*)dst, pix_per_sample, c,
                            cpad, nb));
  return B2U_OK;
}

extern "C" int b2u_threshold_counts(const float* prob, const float* target, long long count,
                                    const float* thresholds, int nthr, double* tp, double* sum_pr, double* sum_gt,
                                    void* stream) {
  B2U_REQUIRE(nthr > 0 && nthr <= 4096, "threshold_counts: 1..4096 thresholds");
  if (count == 0) return B2U_OK;
  int grid = stream_grid(count, kThreads, 4);
  size_t smem = (2 * (size_t)nthr + 1) * sizeof(float);
  B2U_LAUNCH(threshold_counts_kernel, grid, kThreads, smem, stream, prob, target, count, thresholds, nthr, tp,
             sum_pr, sum_gt);
  return B2U_OK;
}

// op lists: with a workspace the large-K Dense(32) runs split-K (accumulator in the workspace)
int b2u_dense_fwd_ws(int dt, const void* x, int k, const float* w, const float* bias, int act, void* y, int m, int n,
                     void* ws, size_t ws_bytes, void* stream) {
  if (m == 32 && k >= 4096 && ws != nullptr && (size_t)n * m * sizeof(float) <= ws_bytes) {
    B2U_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)n * m * sizeof(float), (cudaStream_t)stream));
    dim3 grid(b2u_cdiv(k, kDK), b2u_cdiv(n, kDN));
    DISPATCH_T(dt, B2U_LAUNCH(dense32_splitk_kernel<T>, grid, kThreads, 0, stream, (const T*)x, k, w, n, (float*)ws));
    const long long total = (long long)n * m;
    B2U_LAUNCH(dense_finalize_kernel, (int)((total + kThreads - 1) / kThreads), kThreads, 0, stream, (const float*)ws, bias,
               act, m, total, (float*)y);
    return B2U_OK;
  }
  return b2u_dense_fwd(dt, x, k, w, bias, act, y, m, n, stream);
}

extern "C" int b2u_dense_fwd(int dt, const void* x, int k, const float* w, const float* bias, int act, void* y,
                             int m, int n, void* stream) {
  B2U_REQUIRE(m == 32 || m == 1, "dense_fwd: units must be 32 or 1 (got %d)", m);
  if (m == 32) { DISPATCH_T(dt, B2U_LAUNCH((dense_fwd_kernel<T, 32>), n, kThreads, 0, stream, (const T*)x, k, w, bias, act, (float*)y)); }
  else { DISPATCH_T(dt, B2U_LAUNCH((dense_fwd_kernel<T, 1>), n, kThreads, 0, stream, (const T*)x, k, w, bias, act, (float*)y)); }
  return B2U_OK;
}

extern "C" int b2u_dense_bwd(int dt, const void* x, int k, const float* w, const void* y, int act, const void* dy,
                             void* dx, const void* mask, int mask_act, float* dw, float* db, int m, int n,
                             void* stream) {
  B2U_REQUIRE(m == 32 || m == 1, "dense_bwd: units must be 32 or 1 (got %d)", m);
  B2U_REQUIRE((size_t)n * m * sizeof(float) <= 48 * 1024, "dense_bwd: batch too large for one pass");
  size_t smem = (size_t)n * m * sizeof(float);
  int grid = b2u_cdiv(k, kThreads);
  if (m == 32) { DISPATCH_T(dt, B2U_LAUNCH((dense_bwd_kernel<T, 32>), grid, kThreads, smem, stream, (const T*)x, k, w, (const float*)y, act, (const float*)dy, (T*)dx, (const T*)mask, mask_act, dw, db, n)); }
  else { DISPATCH_T(dt, B2U_LAUNCH((dense_bwd_kernel<T, 1>), grid, kThreads, smem, stream, (const T*)x, k, w, (const float*)y, act, (const float*)dy, (T*)dx, (const T*)mask, mask_act, dw, db, n)); }
  return B2U_OK;
}

extern "C" int b2u_bce_fwd(const float* prob, const float* target, const float* sample_w, int n, float* out,
                           void* stream) {
  B2U_LAUNCH(bce_fwd_kernel, 1, kThreads, 0, stream, prob, target, sample_w, n, out);
  return B2U_OK;
}

extern "C" int b2u_bce_sigmoid_bwd(int dt, const float* prob, const float* target, const float* sample_w, int n,
                                   const b2u_step_state* d_state, void* dlogit, void* stream) {
  DISPATCH_T(dt, B2U_LAUNCH(bce_sigmoid_bwd_kernel<T>, b2u_cdiv(n, 128), 128, 0, stream, prob, target, sample_w, n,
                            d_state, (T*)dlogit));
  return B2U_OK;
}
