// tcgen05 3x3 convolution for THIN layers (Cout <= 64) with the three dw taps side by side in N.
//
// Why: an SS-mode tcgen05.mma reads 128 x 16 fp16 of A (4 KB) per instruction whatever N is; at N = Cout = 32 the
// halo-tile kernel (conv_tc3.cu) is bound by that shared-memory traffic (measured 40-60 cycles per MMA against 16 math
// cycles).  Here a CTA tile is an (8 rows x 16 columns) block of INPUT-aligned pixels: 14 output columns plus one halo
// column on each side.  Accumulator rows = those 128 pixels, columns = (dw, co):
//
//     D[(r, c')][dw*J + co] = sum_dh sum_ci  X[h0 + r + dh - 1][w0 - 1 + c'][ci] * W[dh][dw][ci][co]
//     y[h0 + r][w0 + j][co] = D[(r, j)][0*J + co] + D[(r, j+1)][1*J + co] + D[(r, j+2)][2*J + co]          j = 0..13
//
// so one MMA of N = 3*J replaces three of N = J (A is read a third as often), the three dh taps accumulate into the same
// accumulator (tap row dh = the ordinary K-major descriptor started 16*dh pixel rows further down the contiguous
// (10 x 16)-pixel patch: no halo-pitch trick, every start is swizzle-atom aligned), and the dw shift is two
// __shfl_down_sync per output value in the epilogue (a warp holds two block rows, the neighbours are lanes +1 / +2).
// The packed weights are the halo kernel's [9][J][K] bank read through a (K, 3J, 3) tensor map: [t = dh*3 + dw][j][k]
// IS [dh][dw*J + j][k].
//
// Round-2 GPU history: the first version (validated against the emulator on B200) was epilogue-bound -- three
// tcgen05.ld + wait round trips per 16-column chunk and a shuffle-transpose + shared-memory atomics per chunk for the
// statistics made it 2x slower than the halo kernel on the K = 32 layers.  This version issues the three loads of a chunk
// back to back with ONE wait, keeps BatchNorm statistics / column sums in registers across all tiles of the persistent
// CTA, reads / writes 1-bit ReLU masks, and specialises the epilogue at compile time like conv_tc3.cu.
//
// Warp roles as in conv_tc3.cu: TMA producer, one MMA-issuing warp (elected lane), 4 or 8 epilogue warps, two TMEM
// accumulator stages, weights resident in shared memory for the whole persistent CTA.
#include <cuda.h>
#include "common.cuh"
#include "internal.h"
#include "launch.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kBH = 8, kBW = 16, kOW = kBW - 2;      // block rows / columns, output columns per tile
constexpr int kMaxEpiW = 8;
constexpr int kThreadsW3 = 64 + 32 * kMaxEpiW;
constexpr int kMaxSAw = 8;

struct W3Params {
  int N, H, W, K, J, KS;
  int SA;                 // A ring depth
  int epi_warps;          // 4 (two CTAs per SM) or 8
  uint32_t a_stage;       // bytes of one (10 x 16)-pixel patch slab, 1024-aligned
  uint32_t b_tile;        // bytes of one (dh, slab) weight tile [3J][KS], 1024-aligned
  __half* y; int ldy;
  const float* bias; int act;
  const __half* mask; int ldmask; int mask_act;
  int accumulate;
  double* stats;          // BatchNorm statistics of the stored values: sums at [c], squares at [J + c]
  float* colsum;          // per-channel sums of the stored values (fp32 atomics)
  uint8_t* bits_out;      // kW_BITS_OUT: packed 1-bit ReLU mask of the values stored (bit pix*J + column)
  const uint8_t* bits_in; // kW_BITS_IN : packed 1-bit ReLU mask applied to the values written (data gradient)
  long long* dbg;         // optional timeline buffer (CTA 0): [iter][8] clock64 stamps (b2u_set_option("tc_debug", 1))
};

struct W3Maps {
  CUtensorMap a;          // activations, box (KS, 16, 10, 1)
  CUtensorMap b;          // packed weights [3][3J][K], box (KS, 3J, 1)
};

// lane j ends with the sum over the 32 lanes of v[j] (16 columns); as in conv_tc3.cu
__device__ __forceinline__ float transpose_reduce16w(float v[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
#pragma unroll
    for (int i = 0; i < s; ++i) {
      float a = v[i], b = v[i + s];
      bool up = (lane & s) != 0;
      float send = up ? a : b, keep = up ? b : a;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// epilogue features as compile-time variants (see conv_tc3.cu): bit 0 = fp16 activation mask and/or accumulate,
// bit 1 = BatchNorm statistics and/or column sums, bit 2 = write a 1-bit ReLU mask, bit 3 = read one
constexpr int kW_MASKACC = 1, kW_SUMS = 2, kW_BITS_OUT = 4, kW_BITS_IN = 8, kW_PACKED = 16;

// variants without register statistics are capped at 102 registers: two CTAs of 320 threads then share an SM
template <int kFlags>
__global__ void __launch_bounds__(kThreadsW3, (kFlags & 2) ? 1 : 2) tc_conv3w_kernel(const __grid_constant__ W3Maps maps,
                                                                   const __grid_constant__ W3Params prm) {
  constexpr bool kMaskAcc = (kFlags & kW_MASKACC) != 0, kSums = (kFlags & kW_SUMS) != 0;
  B2U_PDL_LAUNCH_DEPENDENTS();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KS = prm.KS, J = prm.J, NT = 3 * prm.J, SA = prm.SA;
  const int kslabs = prm.K / KS;
  const uint32_t rowb = KS * 2;                                   // bytes per pixel row of a slab
  // layout: [A ring][resident weights 3 x kslabs tiles][barriers][tmem ptr][bias][stats]
  uint8_t* a_ring = smem;
  uint8_t* b_area = a_ring + (size_t)SA * prm.a_stage;
  uint8_t* tail = b_area + (size_t)3 * kslabs * prm.b_tile;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* a_empty = a_full + kMaxSAw;
  uint64_t* w_full = a_empty + kMaxSAw;
  uint64_t* tfull = w_full + 1;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 4) + 15) & ~(uintptr_t)15);
  float* s_stats = s_bias + J;                                    // [2*J]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tiles_w = (prm.W + kOW - 1) / kOW, tiles_h = (prm.H + kBH - 1) / kBH;
  const int ntiles = prm.N * tiles_h * tiles_w;                   // host guarantees < 2^31
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * NT)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    tc::mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], prm.epi_warps); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, tmem_cols);
  B2U_PDL_WAIT();
  for (int i = threadIdx.x; i < J; i += blockDim.x) s_bias[i] = prm.bias ? prm.bias[i] : 0.f;
  for (int i = threadIdx.x; i < 2 * J; i += blockDim.x) s_stats[i] = 0.f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer =========================================
    if (lane == 0) {
      tc::prefetch_tmap(&maps.a);
      tc::prefetch_tmap(&maps.b);
      tc::mbar_expect_tx(w_full, 3u * kslabs * (uint32_t)NT * rowb);
      for (int dh = 0; dh < 3; ++dh)
        for (int ks = 0; ks < kslabs; ++ks)
          tc::tma_load_3d(b_area + (size_t)(dh * kslabs + ks) * prm.b_tile, &maps.b, w_full, ks * KS, 0, dh);
      int sa = 0;
      uint32_t pa = 0;
      const uint32_t a_tx = (uint32_t)(kBH + 2) * kBW * rowb;
      for (unsigned tile = blockIdx.x; tile < (unsigned)ntiles; tile += gridDim.x) {
        const int tw = (int)(tile % (unsigned)tiles_w);
        const unsigned r = tile / (unsigned)tiles_w;
        const int th = (int)(r % (unsigned)tiles_h), n = (int)(r / (unsigned)tiles_h);
        for (int ks = 0; ks < kslabs; ++ks) {
          tc::mbar_wait(&a_empty[sa], pa ^ 1);
          if (prm.dbg && blockIdx.x == 0 && tile / gridDim.x < 64) prm.dbg[(tile / gridDim.x) * 8 + 0] = clock64();
          tc::mbar_expect_tx(&a_full[sa], a_tx);
          // out-of-image rows / columns are zero-filled by TMA = the conv padding
          tc::tma_load_4d(a_ring + (size_t)sa * prm.a_stage, &maps.a, &a_full[sa], ks * KS, tw * kOW - 1, th * kBH - 1, n);
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    const uint32_t idesc = tc::idesc_f16(128, NT, 0, 0);
    const uint64_t layout = KS == 64 ? tc::SWZ_128B : (KS == 32 ? tc::SWZ_64B : tc::SWZ_32B);
    // both operands are plain K-major tiles of contiguous rows: 8-row groups are 8 * rowb apart
    const uint32_t ab_hi = (uint32_t)(tc::smem_desc(0, 16, 8 * rowb, layout) >> 32);
    const uint32_t a_ring_lo = ((tc::smem_u32(a_ring) & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t b_area_lo = ((tc::smem_u32(b_area) & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t a_stage16 = prm.a_stage >> 4, b_tile16 = prm.b_tile >> 4;
    const uint32_t dh_off = ((uint32_t)kBW * rowb) >> 4;          // one block row of 16 pixels, in 16-byte units
    const int ksteps = KS / 16;
    tc::mbar_wait(w_full, 0);
    tc::fence_after_sync();
    int sa = 0, acc = 0;
    uint32_t pa = 0, acc_phase = 0;
    for (unsigned tile = blockIdx.x; tile < (unsigned)ntiles; tile += gridDim.x) {
      tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc::fence_after_sync();
      const bool dbg_on = prm.dbg && blockIdx.x == 0 && tile / gridDim.x < 64;
      if (dbg_on && lane == 0) prm.dbg[(tile / gridDim.x) * 8 + 2] = clock64();
      const uint32_t d_tmem = tmem_base + acc * NT;
      for (int ks = 0; ks < kslabs; ++ks) {
        tc::mbar_wait(&a_full[sa], pa);
        tc::fence_after_sync();
        if (dbg_on && ks == 0 && lane == 0) prm.dbg[(tile / gridDim.x) * 8 + 3] = clock64();
        const uint32_t a_lo = a_ring_lo + (uint32_t)sa * a_stage16;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
          const uint32_t at = a_lo + dh * dh_off;
          const uint32_t bt = b_area_lo + (uint32_t)(dh * kslabs + ks) * b_tile16;
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t ad = ((uint64_t)ab_hi << 32) | (uint64_t)(at + 2 * kk);
            const uint64_t bd = ((uint64_t)ab_hi << 32) | (uint64_t)(bt + 2 * kk);
            tc::mma_f16_ss_elect(d_tmem, ad, bd, idesc, (ks | dh | kk) != 0 ? 1u : 0u);
          }
        }
        tc::mma_commit_elect(&a_empty[sa]);
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
      tc::mma_commit_elect(&tfull[acc]);
      if (dbg_on && lane == 0) prm.dbg[(tile / gridDim.x) * 8 + 4] = clock64();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================== epilogue ==============================================
    // warp -> TMEM lane group (warp & 3); with 8 warps the two warps of a lane group split the output channels.
    // lane group lg holds block rows 2*lg and 2*lg + 1; thread = block pixel (r, c'), it produces output column
    // j = c' (valid for c' <= 13) from its own dw = 0 columns and the dw = 1 / 2 columns of lanes +1 / +2.
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int half = ew >> 2;
    const int blk = lg * 32 + lane;                               // accumulator row = block pixel
    const int r = blk >> 4, cp = blk & 15;
    const bool split = prm.epi_warps == 8 && J % 32 == 0;
    const int ccols = split ? J / 2 : J;                          // output channels owned by this warp
    const int cbeg = split ? half * ccols : 0;
    const bool has_cols = split || half == 0;
    const bool want_sums = kSums && (prm.stats != nullptr || prm.colsum != nullptr);
    // statistics of <= 32 owned columns live in registers across all tiles of the persistent CTA
    const bool reg_stats = want_sums && ccols <= 32;
    constexpr int kRS = kSums ? 32 : 1, kR16 = kSums ? 16 : 0;
    float rs1[kRS], rs2[kRS];
#pragma unroll
    for (int i = 0; i < kRS; ++i) { rs1[i] = 0.f; rs2[i] = 0.f; }
    int acc = 0;
    uint32_t acc_phase = 0;
    // tile coordinates advance incrementally (gridDim.x = d2 * tiles_h * tiles_w + d1 * tiles_w + d0): the three
    // integer divisions per tile were a fifth of this warp's stall samples (ncu source view, round 2)
    const int d0 = (int)(gridDim.x % (unsigned)tiles_w);
    const int d1 = (int)((gridDim.x / (unsigned)tiles_w) % (unsigned)tiles_h);
    const int d2 = (int)(gridDim.x / (unsigned)(tiles_w * tiles_h));
    int tw = (int)(blockIdx.x % (unsigned)tiles_w);
    int th = (int)((blockIdx.x / (unsigned)tiles_w) % (unsigned)tiles_h);
    int n = (int)(blockIdx.x / (unsigned)(tiles_w * tiles_h));
    for (unsigned tile = blockIdx.x; tile < (unsigned)ntiles; tile += gridDim.x) {
      const int h = th * kBH + r, w = tw * kOW + cp;
      const bool valid = cp < kOW && h < prm.H && w < prm.W;
      const long long pix = ((long long)n * prm.H + (h < prm.H ? h : 0)) * prm.W + (w < prm.W ? w : 0);
      __half* yrow = prm.y + pix * prm.ldy;
      const __half* mrow = (kMaskAcc && prm.mask != nullptr) ? prm.mask + pix * prm.ldmask : nullptr;
      // 1-bit ReLU mask of this thread's columns (<= 64): 2-byte loads issued before the accumulator wait
      uint32_t mb0 = 0u, mb1 = 0u;
      if constexpr ((kFlags & kW_BITS_IN) != 0) {
        if (has_cols && valid) {
          const unsigned short* bp16 = reinterpret_cast<const unsigned short*>(prm.bits_in + ((pix * J + cbeg) >> 3));
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q * 16 < ccols) {
              const uint32_t hw = (uint32_t)__ldg(bp16 + q) << (16 * (q & 1));
              if (q < 2) mb0 |= hw; else mb1 |= hw;
            }
          }
        }
      }
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::fence_after_sync();
      const bool dbg_e = prm.dbg && blockIdx.x == 0 && tile / gridDim.x < 64 && threadIdx.x == 64;
      if (dbg_e) prm.dbg[(tile / gridDim.x) * 8 + 5] = clock64();
      if (has_cols) {
#pragma unroll 1
        for (int cc = 0; cc < ccols; cc += 16) {
          const int c0 = cbeg + cc;
          const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * NT + c0;
          uint32_t a0[16], a1[16], a2[16];
          tc::tmem_ld16_nowait(trow, a0);                          // dw = 0: own pixel
          tc::tmem_ld16_nowait(trow + J, a1);                      // dw = 1: the pixel one column to the right
          tc::tmem_ld16_nowait(trow + 2 * J, a2);                  // dw = 2: two columns to the right
          tc::tmem_wait_ld();
          tc::reg_fence16(a0);
          tc::reg_fence16(a1);
          tc::reg_fence16(a2);
          if (dbg_e && cc == 0) prm.dbg[(tile / gridDim.x) * 8 + 7] = clock64();
          float v[16];
          if constexpr ((kFlags & kW_PACKED) != 0) {
            // training-mode ops: the two shifted partial sums travel as fp16 pairs -- 16 shuffles per chunk instead of 32
            // (the epilogue is bound by the SHFL rate of the SM's MIO pipe: ncu source view, profiles/).  The partials
            // are rounded to fp16 once more than the stored result is; inference ops keep the exact fp32 shuffles.
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __half2 h1 = __floats2half2_rn(__uint_as_float(a1[2 * j]), __uint_as_float(a1[2 * j + 1]));
              __half2 h2 = __floats2half2_rn(__uint_as_float(a2[2 * j]), __uint_as_float(a2[2 * j + 1]));
              const unsigned u1 = __shfl_down_sync(0xffffffffu, *reinterpret_cast<unsigned*>(&h1), 1);
              const unsigned u2 = __shfl_down_sync(0xffffffffu, *reinterpret_cast<unsigned*>(&h2), 2);
              const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u1));
              const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u2));
              v[2 * j] = (__uint_as_float(a0[2 * j]) + f1.x) + f2.x;
              v[2 * j + 1] = (__uint_as_float(a0[2 * j + 1]) + f1.y) + f2.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              v[i] = (__uint_as_float(a0[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(a1[i]), 1)) +
                     __shfl_down_sync(0xffffffffu, __uint_as_float(a2[i]), 2);
          }
          const float4* bp = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = bp[q];
            v[4 * q + 0] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
          }
          if (prm.act == B2U_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (prm.act == B2U_ACT_ELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
          }
          if constexpr ((kFlags & kW_BITS_IN) != 0) {
            const uint32_t w32 = cc < 32 ? mb0 : mb1;
            const uint32_t b16 = (w32 >> (cc & 16)) & 0xffffu;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ((b16 >> i) & 1u) ? v[i] : 0.f;
          }
          if (valid) {
            if constexpr (kMaskAcc) {
              if (mrow != nullptr) {
                float m[8];
                load8<__half>(mrow + c0, m);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= act_bwd_from_y(m[i], prm.mask_act);
                load8<__half>(mrow + c0 + 8, m);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[8 + i] *= act_bwd_from_y(m[i], prm.mask_act);
              }
              if (prm.accumulate) {
                float e[8];
                load8<__half>(yrow + c0, e);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += e[i];
                load8<__half>(yrow + c0 + 8, e);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[8 + i] += e[i];
              }
            }
            store8<__half>(yrow + c0, v);
            store8<__half>(yrow + c0 + 8, v + 8);
            if constexpr ((kFlags & kW_BITS_OUT) != 0) {
              // bit = (the fp16 value just stored > 0): round-to-nearest-even sends v <= 2^-25 to zero
              uint32_t b16 = 0u;
#pragma unroll
              for (int i = 0; i < 16; ++i) b16 |= (v[i] > 2.98023223876953125e-08f ? 1u : 0u) << i;
              *reinterpret_cast<unsigned short*>(prm.bits_out + ((pix * J + c0) >> 3)) = (unsigned short)b16;
            }
          }
          if (want_sums) {
            if (reg_stats) {
              if (valid) {
                if (cc == 0) {
#pragma unroll
                  for (int i = 0; i < kR16; ++i) { rs1[i] += v[i]; rs2[i] = fmaf(v[i], v[i], rs2[i]); }
                } else {
#pragma unroll
                  for (int i = 0; i < kR16; ++i) { rs1[kR16 + i] += v[i]; rs2[kR16 + i] = fmaf(v[i], v[i], rs2[kR16 + i]); }
                }
              }
            } else {
              float q[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) q[i] = valid ? v[i] : 0.f;
              if (prm.stats != nullptr) {
                float sq[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) sq[i] = q[i] * q[i];
                const float s2 = transpose_reduce16w(sq, lane);
                if (lane < 16) atomicAdd(&s_stats[J + c0 + lane], s2);
              }
              const float s1 = transpose_reduce16w(q, lane);
              if (lane < 16) atomicAdd(&s_stats[c0 + lane], s1);
            }
          }
        }
      }
      if (dbg_e) prm.dbg[(tile / gridDim.x) * 8 + 6] = clock64();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      tw += d0;
      th += d1;
      n += d2;
      if (tw >= tiles_w) { tw -= tiles_w; ++th; }
      if (th >= tiles_h) { th -= tiles_h; ++n; }
    }
    if (reg_stats && has_cols) {
      for (int cc = 0; cc < ccols; cc += 16) {
        float q[16], sq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int i0 = kSums ? i : 0, i1 = kSums ? 16 + i : 0;       // (never executed without kSums)
          q[i] = cc == 0 ? rs1[i0] : rs1[i1];
          sq[i] = cc == 0 ? rs2[i0] : rs2[i1];
        }
        const float s1 = transpose_reduce16w(q, lane);
        if (lane < 16) atomicAdd(&s_stats[cbeg + cc + lane], s1);
        if (prm.stats != nullptr) {
          const float s2 = transpose_reduce16w(sq, lane);
          if (lane < 16) atomicAdd(&s_stats[J + cbeg + cc + lane], s2);
        }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (prm.stats != nullptr) {
    for (int i = threadIdx.x; i < 2 * J; i += blockDim.x) {
      const float s = s_stats[i];
      if (s != 0.f) atomicAdd(&prm.stats[i], (double)s);
    }
  }
  if (prm.colsum != nullptr) {
    for (int i = threadIdx.x; i < J; i += blockDim.x) {
      const float s = s_stats[i];
      if (s != 0.f) atomicAdd(&prm.colsum[i], s);
    }
  }
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// fwd: Wp[dh][dw*J + co][ci] = w[dh*3 + dw][ci][co]     (w: Keras HWIO, K = Cin, J = Cout)
// dgrad: Wp[dh][dw*J + j][k] = w[8 - (dh*3 + dw)][j][k]  (j = Cin of the forward = columns written, k = its Cout)
// (the same bytes as conv_tc3.cu's [9][J][K] bank: used only when the caller has no prepacked copy)
__global__ void pack3w_kernel(const float* __restrict__ w, __half* __restrict__ wp, int dgrad, int J, int K) {
  B2U_PDL_PROLOGUE();
  const long long total = 9LL * J * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long r = i / K;
    const int nrow = (int)(r % (3 * J)), dh = (int)(r / (3 * J));
    const int dw = nrow / J, j = nrow % J;
    const int t = dh * 3 + dw;
    const float v = dgrad ? w[((long long)(8 - t) * J + j) * K + k] : w[((long long)t * K + k) * J + j];
    wp[i] = __float2half_rn(v);
  }
}

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnW g_encw = nullptr;
bool g_attrw = false;

int get_encw() {
  if (g_encw != nullptr) return B2U_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B2U_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    b2u_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B2U_ERR_CUDA;
  }
  g_encw = (EncodeTiledFnW)fn;
  return B2U_OK;
}

}  // namespace

// 0: never, 1: every layer the kernel takes (A/B runs), 2: the layers where it measured faster than the halo kernel
int g_b2u_tc_dwmerge = 2;
int g_b2u_tc_dw_epi8 = 1;      // 8 epilogue warps also in the two-CTAs-per-SM configuration (variants without statistics)

// shapes this kernel takes: resident weights and the A ring must fit shared memory
int b2u_tc_conv3x3_dwmerge_ok(int K, int J) {
  if (J % 16 || J > 64 || J < 16) return 0;
  const int KS = K % 64 == 0 ? 64 : (K % 32 == 0 ? 32 : (K % 16 == 0 ? 16 : 0));
  if (KS == 0) return 0;
  const size_t rowb = KS * 2;
  const size_t a_stage = ((size_t)(kBH + 2) * kBW * rowb + 1023) & ~(size_t)1023;
  const size_t b_tile = ((size_t)3 * J * rowb + 1023) & ~(size_t)1023;
  const size_t wres = 3 * (size_t)(K / KS) * b_tile;
  return wres + 2 * a_stage + 4096 <= 200 * 1024;
}

int b2u_tc_conv3x3_dwmerge(const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act, void* y,
                           int ldy, int J, double* stats, float* colsum, const void* mask, int ldmask, int mask_act,
                           int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp, void* stream,
                           void* relu_bits_out, int packed_shift) {
  int rc = get_encw();
  if (rc != B2U_OK) return rc;
  B2U_REQUIRE(b2u_tc_conv3x3_dwmerge_ok(K, J), "tc_conv3w: unsupported channel counts K=%d J=%d", K, J);
  W3Params p{};
  p.N = n; p.H = h; p.W = wd; p.K = K; p.J = J;
  p.KS = K % 64 == 0 ? 64 : (K % 32 == 0 ? 32 : 16);
  p.y = (__half*)y; p.ldy = ldy; p.bias = bias; p.act = act;
  p.mask = (const __half*)mask; p.ldmask = ldmask; p.mask_act = mask_act; p.accumulate = accumulate;
  p.stats = stats; p.colsum = colsum;
  p.bits_out = (uint8_t*)relu_bits_out;
  p.dbg = g_b2u_dbg;
  if (mask != nullptr && mask_act == B2U_ACT_RELU_BITS) {          // packed 1-bit mask instead of the activation tensor
    p.bits_in = (const uint8_t*)mask;
    p.mask = nullptr;
    mask = nullptr;
  }
  const uint32_t rowb = p.KS * 2;
  const int kslabs = K / p.KS, NT = 3 * J;
  p.a_stage = (uint32_t)(((size_t)(kBH + 2) * kBW * rowb + 1023) & ~(size_t)1023);
  p.b_tile = (uint32_t)(((size_t)NT * rowb + 1023) & ~(size_t)1023);
  const size_t wres = 3 * (size_t)kslabs * p.b_tile;
  const size_t tail = (2 * kMaxSAw + 5) * 8 + 32 + (size_t)3 * J * 4 + 64;
  // two CTAs per SM when TMEM (2 * 3J columns each, power of two) and shared memory allow it
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * NT)) cols <<= 1;
  bool two = cols <= 256 && 1024 + wres + tail + 2 * (size_t)p.a_stage <= 110 * 1024;
  const size_t cap = two ? 110 * 1024 : 222 * 1024;
  p.SA = (int)((cap - 1024 - wres - tail) / p.a_stage);
  if (p.SA > kMaxSAw) p.SA = kMaxSAw;
  B2U_REQUIRE(p.SA >= 2, "tc_conv3w: tiles do not fit shared memory (K=%d J=%d)", K, J);
  // two CTAs per SM: 4 epilogue warps each -- or 8 for the variants without register statistics (96 registers:
  // 2 x 320 threads fit the register file), whose epilogue is bound by the latency of its shuffles (ncu: the FADDs
  // behind the SHFLs hold most stall samples), so twice the warps hide twice the latency
  const bool has_sums = stats != nullptr || colsum != nullptr;
  p.epi_warps = (two && !(g_b2u_tc_dw_epi8 && !has_sums)) ? 4 : 8;
  const size_t smem = 1024 + (size_t)p.SA * p.a_stage + wres + tail;

  if (wp == nullptr) {                     // no prepacked bank from the caller: pack into the workspace
    const size_t need = 9 * (size_t)J * K * 2;
    B2U_REQUIRE(ws != nullptr && need <= ws_bytes, "tc_conv3w: workspace too small");
    const long long total = 9LL * J * K;
    int grid = (int)((total + 255) / 256);
    if (grid > 8 * B2U_NUM_SMS) grid = 8 * B2U_NUM_SMS;
    B2U_LAUNCH(pack3w_kernel, grid, 256, 0, stream, w, (__half*)ws, dgrad, J, K);
    wp = ws;
  }
  W3Maps maps;
  {
    cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)wd * ldx * 2, (cuuint64_t)h * wd * ldx * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.KS, (cuuint32_t)kBW, (cuuint32_t)(kBH + 2), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = p.KS == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (p.KS == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = g_encw(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b2u_set_error("tc_conv3w: activation tensor map failed (%d)", (int)r); return B2U_ERR_CUDA; }
    cuuint64_t bd[3] = {(cuuint64_t)K, (cuuint64_t)NT, 3};
    cuuint64_t bs[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * NT * 2};
    cuuint32_t bb[3] = {(cuuint32_t)p.KS, (cuuint32_t)NT, 1};
    cuuint32_t be[3] = {1, 1, 1};
    r = g_encw(&maps.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(wp), bd, bs, bb, be,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b2u_set_error("tc_conv3w: weight tensor map failed (%d)", (int)r); return B2U_ERR_CUDA; }
  }
  if (!g_attrw) {
#define B2U_W3_ATTR(F)                                                                                                   \
  B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3w_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));     \
  B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3w_kernel<F | kW_PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    B2U_W3_ATTR(0); B2U_W3_ATTR(1); B2U_W3_ATTR(2); B2U_W3_ATTR(3); B2U_W3_ATTR(4); B2U_W3_ATTR(6); B2U_W3_ATTR(8); B2U_W3_ATTR(10);
#undef B2U_W3_ATTR
    g_attrw = true;
  }
  const long long tiles = (long long)n * b2u_cdiv(h, kBH) * b2u_cdiv(wd, kOW);
  B2U_REQUIRE(tiles < (1LL << 31), "tc_conv3w: too many tiles");
  int ctas = B2U_NUM_SMS;
  if (two && tiles >= 4 * B2U_NUM_SMS) ctas = 2 * B2U_NUM_SMS;
  const int grid = (int)(tiles < ctas ? tiles : ctas);
  const int flags = ((mask != nullptr || accumulate) ? kW_MASKACC : 0) | ((stats != nullptr || colsum != nullptr) ? kW_SUMS : 0) |
                    (p.bits_out != nullptr ? kW_BITS_OUT : 0) | (p.bits_in != nullptr ? kW_BITS_IN : 0);
  const int nthr = 64 + 32 * p.epi_warps;
#define B2U_W3_CASE(F)                                                                                     \
  case F:                                                                                                  \
    if (packed_shift) B2U_LAUNCH(tc_conv3w_kernel<F | kW_PACKED>, grid, nthr, smem, stream, maps, p);      \
    else B2U_LAUNCH(tc_conv3w_kernel<F>, grid, nthr, smem, stream, maps, p);                               \
    break
  switch (flags) {
    B2U_W3_CASE(0); B2U_W3_CASE(1); B2U_W3_CASE(2); B2U_W3_CASE(3);
    B2U_W3_CASE(4);        // forward + bit mask
    B2U_W3_CASE(6);        // ... + statistics
    B2U_W3_CASE(8);        // data gradient, 1-bit mask
    B2U_W3_CASE(10);       // ... + column sums
    default:
      b2u_set_error("tc_conv3w: unsupported feature combination %d (1-bit masks do not combine with accumulate / fp16 masks)", flags);
      return B2U_ERR_ARG;
  }
#undef B2U_W3_CASE
  return B2U_OK;
}
