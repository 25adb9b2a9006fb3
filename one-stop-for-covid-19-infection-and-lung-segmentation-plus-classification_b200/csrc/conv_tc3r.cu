// tcgen05 3x3 convolution for the THINNEST layers (Cout = 16 / 32): row strips with the three dh taps merged in N and a
// shuffle-free epilogue.
//
// Where the other two kernels stand (B200, round 2, K = N = 32 at 512^2): the halo-tile kernel (conv_tc3.cu) issues nine
// MMAs of N = 32 per K step and is bound by the 4 KB A-operand read every SS-mode tcgen05.mma makes whatever N is
// (~62 clk per MMA, 9.1 clk per pixel and SM against an HBM floor of 5.5); the dw-merged kernel (conv_tc3w.cu) needs a
// third of the MMAs (N = 96) but must shift two of its three partial sums across LANES, and the SM's SHFL rate then
// bounds its epilogue.  This kernel merges the dh taps instead and lays the accumulator out so that the three partial
// sums of an output pixel sit in the SAME TMEM lane:
//
//   * accumulator rows = 128 consecutive pixels of ONE image row r (a "row strip"), columns = (dh, co):
//         Acc_r[w][dh*J + co] = sum_dw sum_ci  X[r][w + dw - 1][ci] * W[dh][dw][ci][co]
//     one TMA box (KS, 130 pixels, 1 row) per input row and K slab; the dw shift is the A descriptor started dw pixel
//     rows further (contiguous K-major rows, as in the halo kernel); 3 * K/16 MMAs of N = 3J per input row;
//   * output row h needs input rows h-1, h, h+1:   y[h][w][co] = Acc_{h-1}[w][0*J+co] + Acc_h[w][1*J+co] + Acc_{h+1}[w][2*J+co]
//     -- three accumulators of a ring of five in TMEM (5 x 96 columns), all read by the thread that owns lane w: three
//     tcgen05.ld, two FADDs per value, no shuffle;
//   * a persistent CTA owns a contiguous range of (image, strip, row) units, so every input row is loaded and
//     multiplied once (plus one halo row at each end of the range: 2 of ~55 rows).
//
// Warp roles: TMA producer, one MMA-issuing warp (elected lane), 8 epilogue warps (TMEM lane quadrant x column half).
// One CTA per SM (the accumulator ring takes the whole TMEM).  Weights stay resident in shared memory.
#include <cuda.h>
#include "common.cuh"
#include "internal.h"
#include "launch.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kSW = 128;                   // pixels per strip (accumulator rows)
constexpr int kPatch = kSW + 2;            // input pixels per row box
constexpr int kNS = 5;                     // accumulator ring depth
constexpr int kThreadsR = 64 + 32 * 8;
constexpr int kMaxSAr = 8;

struct R3Params {
  int N, H, W, K, J, KS;
  int SA;                 // A ring depth
  int strips;             // ceil(W / 128)
  uint32_t a_stage;       // bytes of one (130-pixel) row slab, 1024-aligned
  uint32_t b_tile;        // bytes of one (dw, slab) weight tile [3J][KS], 1024-aligned
  __half* y; int ldy;
  const float* bias; int act;
  const __half* mask; int ldmask; int mask_act;
  int accumulate;
  double* stats;          // BatchNorm statistics of the stored values: sums at [c], squares at [J + c]
  float* colsum;          // per-channel sums of the stored values (fp32 atomics)
  uint8_t* bits_out;      // packed 1-bit ReLU mask of the values stored (bit pix*J + column)
  const uint8_t* bits_in; // packed 1-bit ReLU mask applied to the values written (data gradient)
  long long* dbg;         // optional timeline buffer (CTA 0): [row][8] clock64 stamps (b2u_set_option("tc_debug", 1))
};

struct R3Maps {
  CUtensorMap a;          // activations (K, W, H, N), box (KS, 130, 1, 1)
  CUtensorMap b;          // packed weights viewed as (K, J, dw, dh), box (KS, J, 1, 3)
};

__device__ __forceinline__ float transpose_reduce16r(float v[16], int lane) {
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

constexpr int kR_MASKACC = 1, kR_SUMS = 2, kR_BITS_OUT = 4, kR_BITS_IN = 8;

// segment iteration shared by the three roles: the CTA's unit range [u, u1) cut at (image, strip) boundaries
struct Seg {
  int n, strip, ha, hb;
};
__device__ __forceinline__ Seg next_seg(long long& u, long long u1, int H, int strips) {
  Seg s;
  const long long q = u / H;
  s.ha = (int)(u - q * H);
  s.strip = (int)(q % strips);
  s.n = (int)(q / strips);
  const long long room = u1 - u;
  s.hb = (long long)(H - s.ha) < room ? H : s.ha + (int)room;
  u += s.hb - s.ha;
  return s;
}

template <int kFlags>
__global__ void __launch_bounds__(kThreadsR, 1) tc_conv3r_kernel(const __grid_constant__ R3Maps maps,
                                                                  const __grid_constant__ R3Params prm) {
  constexpr bool kMaskAcc = (kFlags & kR_MASKACC) != 0, kSums = (kFlags & kR_SUMS) != 0;
  B2U_PDL_LAUNCH_DEPENDENTS();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KS = prm.KS, J = prm.J, NT = 3 * prm.J, SA = prm.SA;
  const int kslabs = prm.K / KS;
  const uint32_t rowb = KS * 2;
  uint8_t* a_ring = smem;
  uint8_t* b_area = a_ring + (size_t)SA * prm.a_stage;
  uint8_t* tail = b_area + (size_t)3 * kslabs * prm.b_tile;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* a_empty = a_full + kMaxSAr;
  uint64_t* w_full = a_empty + kMaxSAr;
  uint64_t* acc_full = w_full + 1;
  uint64_t* acc_empty = acc_full + kNS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + kNS);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 4) + 15) & ~(uintptr_t)15);
  float* s_stats = s_bias + J;                                    // [2*J]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int epi_active = J >= 32 ? 8 : 4;                         // J = 16: the second column half has nothing to do
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(kNS * NT)) tmem_cols <<= 1;

  // this CTA's contiguous range of (image, strip, row) units
  const long long units = (long long)prm.N * prm.strips * prm.H;
  const long long ub = units * blockIdx.x / gridDim.x, ue = units * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    tc::mbar_init(w_full, 1);
    for (int s = 0; s < kNS; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], epi_active); }
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
    if (lane == 0 && ub < ue) {
      tc::prefetch_tmap(&maps.a);
      tc::prefetch_tmap(&maps.b);
      tc::mbar_expect_tx(w_full, 3u * kslabs * (uint32_t)NT * rowb);
      for (int dw = 0; dw < 3; ++dw)
        for (int ks = 0; ks < kslabs; ++ks)
          tc::tma_load_4d(b_area + (size_t)(dw * kslabs + ks) * prm.b_tile, &maps.b, w_full, ks * KS, 0, dw, 0);
      int sa = 0;
      uint32_t pa = 0;
      const uint32_t a_tx = (uint32_t)kPatch * rowb;
      long long u = ub;
      int drow = 0;
      while (u < ue) {
        const Seg sg = next_seg(u, ue, prm.H, prm.strips);
        for (int r = sg.ha - 1; r <= sg.hb; ++r, ++drow) {         // input rows incl. one halo row at each end
          for (int ks = 0; ks < kslabs; ++ks) {
            tc::mbar_wait(&a_empty[sa], pa ^ 1);
            if (prm.dbg && blockIdx.x == 0 && drow < 64 && ks == 0) prm.dbg[drow * 8 + 0] = clock64();
            tc::mbar_expect_tx(&a_full[sa], a_tx);
            // out-of-image rows / columns are zero-filled by TMA = the conv padding
            tc::tma_load_4d(a_ring + (size_t)sa * prm.a_stage, &maps.a, &a_full[sa], ks * KS, sg.strip * kSW - 1, r, sg.n);
            if (++sa == SA) { sa = 0; pa ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    if (ub < ue) {
      const uint32_t idesc = tc::idesc_f16(128, NT, 0, 0);
      const uint64_t layout = KS == 64 ? tc::SWZ_128B : (KS == 32 ? tc::SWZ_64B : tc::SWZ_32B);
      const uint32_t ab_hi = (uint32_t)(tc::smem_desc(0, 16, 8 * rowb, layout) >> 32);
      const uint32_t a_ring_lo = ((tc::smem_u32(a_ring) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t b_area_lo = ((tc::smem_u32(b_area) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t a_stage16 = prm.a_stage >> 4, b_tile16 = prm.b_tile >> 4;
      const uint32_t dw_off = rowb >> 4;                            // one pixel row, in 16-byte units
      const int ksteps = KS / 16;
      tc::mbar_wait(w_full, 0);
      tc::fence_after_sync();
      int sa = 0, slot = 0;
      uint32_t pa = 0, slot_phase = 0;
      long long u = ub;
      int drow = 0;
      while (u < ue) {
        const Seg sg = next_seg(u, ue, prm.H, prm.strips);
        const int rows = sg.hb - sg.ha + 2;
        for (int i = 0; i < rows; ++i, ++drow) {
          const bool dbg_on = prm.dbg && blockIdx.x == 0 && drow < 64 && lane == 0;
          if (dbg_on) prm.dbg[drow * 8 + 1] = clock64();
          tc::mbar_wait(&acc_empty[slot], slot_phase ^ 1);
          tc::fence_after_sync();
          if (dbg_on) prm.dbg[drow * 8 + 2] = clock64();
          const uint32_t d_tmem = tmem_base + slot * NT;
          for (int ks = 0; ks < kslabs; ++ks) {
            tc::mbar_wait(&a_full[sa], pa);
            tc::fence_after_sync();
            if (dbg_on && ks == 0) prm.dbg[drow * 8 + 3] = clock64();
            const uint32_t a_lo = a_ring_lo + (uint32_t)sa * a_stage16;
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
              const uint32_t at = a_lo + dw * dw_off;
              const uint32_t bt = b_area_lo + (uint32_t)(dw * kslabs + ks) * b_tile16;
              for (int kk = 0; kk < ksteps; ++kk) {
                const uint64_t ad = ((uint64_t)ab_hi << 32) | (uint64_t)(at + 2 * kk);
                const uint64_t bd = ((uint64_t)ab_hi << 32) | (uint64_t)(bt + 2 * kk);
                tc::mma_f16_ss_elect(d_tmem, ad, bd, idesc, (ks | dw | kk) != 0 ? 1u : 0u);
              }
            }
            tc::mma_commit_elect(&a_empty[sa]);
            if (++sa == SA) { sa = 0; pa ^= 1; }
          }
          tc::mma_commit_elect(&acc_full[slot]);
          if (dbg_on) prm.dbg[drow * 8 + 4] = clock64();
          if (++slot == kNS) { slot = 0; slot_phase ^= 1; }
        }
      }
    }
  } else {
    // ===================================== epilogue ==============================================
    // warp -> TMEM lane quadrant (warp & 3) = pixels [32q, 32q + 32) of the strip, column half ((warp - 2) >> 2)
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    const bool active = half == 0 || J >= 32;
    const int c0 = half * 16;                                      // this warp's 16 output channels
    const bool want_sums = kSums && (prm.stats != nullptr || prm.colsum != nullptr);
    constexpr int kR16 = kSums ? 16 : 1;
    float rs1[kR16], rs2[kR16];
#pragma unroll
    for (int i = 0; i < kR16; ++i) { rs1[i] = 0.f; rs2[i] = 0.f; }
    if (active && ub < ue) {
      long long g = 0;                                             // input rows consumed so far (ring position)
      long long u = ub;
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
      while (u < ue) {
        const Seg sg = next_seg(u, ue, prm.H, prm.strips);
        const int R = sg.hb - sg.ha;
        const int w = sg.strip * kSW + q * 32 + lane;
        const bool valid = w < prm.W;
        auto wait_row = [&](long long gg) {
          tc::mbar_wait(&acc_full[(int)(gg % kNS)], (uint32_t)((gg / kNS) & 1));
        };
        // Software pipeline over the output rows (round-2 measurement: with one row in flight per warp the row period
        // was 1300 clk -- barrier wait + three TMEM loads + arithmetic + stores in series -- against ~600 clk of MMAs):
        // the TMEM loads of row j+1 are issued before the arithmetic of row j, in a second register set.
        auto issue = [&](int jj, uint32_t (&r0)[16], uint32_t (&r1)[16], uint32_t (&r2)[16]) {
          tc::fence_after_sync();
          tc::tmem_ld16_nowait(lane_base + (uint32_t)((g + jj) % kNS) * NT + c0, r0);               // dh = 0: row h-1
          tc::tmem_ld16_nowait(lane_base + (uint32_t)((g + jj + 1) % kNS) * NT + J + c0, r1);       // dh = 1: row h
          tc::tmem_ld16_nowait(lane_base + (uint32_t)((g + jj + 2) % kNS) * NT + 2 * J + c0, r2);   // dh = 2: row h+1
        };
        auto landed = [&](int jj, uint32_t (&r0)[16], uint32_t (&r1)[16], uint32_t (&r2)[16]) {
          tc::tmem_wait_ld();
          tc::reg_fence16(r0);
          tc::reg_fence16(r1);
          tc::reg_fence16(r2);
          // input row g + jj has had its last reader (as dh = 0): hand its accumulator back to the MMA warp
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&acc_empty[(int)((g + jj) % kNS)]);
        };
        auto finish = [&](int jj, const uint32_t (&a0)[16], const uint32_t (&a1)[16], const uint32_t (&a2)[16]) {
          const int h = sg.ha + jj;
          const long long pix = ((long long)sg.n * prm.H + h) * prm.W + (valid ? w : 0);
          __half* yrow = prm.y + pix * prm.ldy + c0;
          uint32_t b16in = 0xffffu;
          if constexpr ((kFlags & kR_BITS_IN) != 0) {
            if (valid) b16in = __ldg(reinterpret_cast<const unsigned short*>(prm.bits_in + ((pix * J + c0) >> 3)));
          }
          float v[16];
          const float4* bp = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const float4 b4 = bp[qd];
            v[4 * qd + 0] = ((__uint_as_float(a0[4 * qd + 0]) + __uint_as_float(a1[4 * qd + 0])) + __uint_as_float(a2[4 * qd + 0])) + b4.x;
            v[4 * qd + 1] = ((__uint_as_float(a0[4 * qd + 1]) + __uint_as_float(a1[4 * qd + 1])) + __uint_as_float(a2[4 * qd + 1])) + b4.y;
            v[4 * qd + 2] = ((__uint_as_float(a0[4 * qd + 2]) + __uint_as_float(a1[4 * qd + 2])) + __uint_as_float(a2[4 * qd + 2])) + b4.z;
            v[4 * qd + 3] = ((__uint_as_float(a0[4 * qd + 3]) + __uint_as_float(a1[4 * qd + 3])) + __uint_as_float(a2[4 * qd + 3])) + b4.w;
          }
          if (prm.act == B2U_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (prm.act == B2U_ACT_ELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
          }
          if constexpr ((kFlags & kR_BITS_IN) != 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ((b16in >> i) & 1u) ? v[i] : 0.f;
          }
          if (valid) {
            if constexpr (kMaskAcc) {
              if (prm.mask != nullptr) {
                const __half* mrow = prm.mask + pix * prm.ldmask + c0;
                float m[8];
                load8<__half>(mrow, m);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= act_bwd_from_y(m[i], prm.mask_act);
                load8<__half>(mrow + 8, m);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[8 + i] *= act_bwd_from_y(m[i], prm.mask_act);
              }
              if (prm.accumulate) {
                float e[8];
                load8<__half>(yrow, e);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += e[i];
                load8<__half>(yrow + 8, e);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[8 + i] += e[i];
              }
            }
            store8<__half>(yrow, v);
            store8<__half>(yrow + 8, v + 8);
            if constexpr ((kFlags & kR_BITS_OUT) != 0) {
              uint32_t b16 = 0u;
#pragma unroll
              for (int i = 0; i < 16; ++i) b16 |= (v[i] > 2.98023223876953125e-08f ? 1u : 0u) << i;
              *reinterpret_cast<unsigned short*>(prm.bits_out + ((pix * J + c0) >> 3)) = (unsigned short)b16;
            }
            if (want_sums) {
#pragma unroll
              for (int i = 0; i < kR16; ++i) { rs1[i] += v[i]; rs2[i] = fmaf(v[i], v[i], rs2[i]); }
            }
          }
        };
        uint32_t pa0[16], pa1[16], pa2[16], pb0[16], pb1[16], pb2[16];
        wait_row(g);
        wait_row(g + 1);
        wait_row(g + 2);
        issue(0, pa0, pa1, pa2);
        for (int j = 0; j < R; j += 2) {
          const bool dbg_e = prm.dbg && blockIdx.x == 0 && threadIdx.x == 64 && g + j < 64;
          if (dbg_e) prm.dbg[(g + j) * 8 + 5] = clock64();
          landed(j, pa0, pa1, pa2);
          if (dbg_e) prm.dbg[(g + j) * 8 + 6] = clock64();
          if (j + 1 < R) {
            wait_row(g + j + 3);
            issue(j + 1, pb0, pb1, pb2);
          }
          if (dbg_e) prm.dbg[(g + j) * 8 + 7] = clock64();
          finish(j, pa0, pa1, pa2);
          if (j + 1 < R) {
            landed(j + 1, pb0, pb1, pb2);
            if (j + 2 < R) {
              wait_row(g + j + 4);
              issue(j + 2, pa0, pa1, pa2);
            }
            finish(j + 1, pb0, pb1, pb2);
          }
        }
        // the two halo rows at the end of the segment: no output row retires them
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(&acc_empty[(int)((g + R) % kNS)]);
          tc::mbar_arrive(&acc_empty[(int)((g + R + 1) % kNS)]);
        }
        g += R + 2;
      }
      if (want_sums) {
        float qv[16], sq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { qv[i] = rs1[kSums ? i : 0]; sq[i] = rs2[kSums ? i : 0]; }
        const float s1 = transpose_reduce16r(qv, lane);
        if (lane < 16) atomicAdd(&s_stats[c0 + lane], s1);
        if (prm.stats != nullptr) {
          const float s2 = transpose_reduce16r(sq, lane);
          if (lane < 16) atomicAdd(&s_stats[J + c0 + lane], s2);
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

// [t][j][k] fp16 bank from the fp32 Keras kernel, for callers without a prepacked copy (same bytes as conv_tc3.cu's)
__global__ void pack3r_kernel(const float* __restrict__ w, __half* __restrict__ wp, int dgrad, int J, int K) {
  B2U_PDL_PROLOGUE();
  const long long total = 9LL * J * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long r = i / K;
    const int j = (int)(r % J), t = (int)(r / J);
    const float v = dgrad ? w[((long long)(8 - t) * J + j) * K + k] : w[((long long)t * K + k) * J + j];
    wp[i] = __float2half_rn(v);
  }
}

typedef CUresult (*EncodeTiledFnR)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnR g_encr = nullptr;
bool g_attrr = false;

int get_encr() {
  if (g_encr != nullptr) return B2U_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B2U_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    b2u_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B2U_ERR_CUDA;
  }
  g_encr = (EncodeTiledFnR)fn;
  return B2U_OK;
}

}  // namespace

// 0: never, 1: every layer the kernel takes, 2 (default): where it measured faster than the halo / dw-merged kernels
int g_b2u_tc_rowstrip = 2;

int b2u_tc_conv3x3_rowstrip_ok(int K, int J, int wd) {
  if (J != 16 && J != 32) return 0;
  const int KS = K % 64 == 0 ? 64 : (K % 32 == 0 ? 32 : (K % 16 == 0 ? 16 : 0));
  if (KS == 0 || wd < 32) return 0;
  const size_t rowb = KS * 2;
  const size_t a_stage = ((size_t)kPatch * rowb + 1023) & ~(size_t)1023;
  const size_t b_tile = ((size_t)3 * J * rowb + 1023) & ~(size_t)1023;
  const size_t wres = 3 * (size_t)(K / KS) * b_tile;
  return wres + 3 * a_stage + 4096 <= 210 * 1024;
}

// `wp`: the prepacked [9][J][K] fp16 bank (the planner's OP_PACK_WEIGHTS output); without it the kernel `w` is packed
// into the workspace first
int b2u_tc_conv3x3_rowstrip(const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act, void* y,
                            int ldy, int J, double* stats, float* colsum, const void* mask, int ldmask, int mask_act,
                            int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp, void* stream,
                            void* relu_bits_out) {
  int rc = get_encr();
  if (rc != B2U_OK) return rc;
  B2U_REQUIRE(b2u_tc_conv3x3_rowstrip_ok(K, J, wd), "tc_conv3r: unsupported shape K=%d J=%d W=%d", K, J, wd);
  if (wp == nullptr) {
    const size_t need = 9 * (size_t)J * K * 2;
    B2U_REQUIRE(w != nullptr && ws != nullptr && need <= ws_bytes, "tc_conv3r: workspace too small");
    const long long total = 9LL * J * K;
    int pgrid = (int)((total + 255) / 256);
    if (pgrid > 8 * B2U_NUM_SMS) pgrid = 8 * B2U_NUM_SMS;
    B2U_LAUNCH(pack3r_kernel, pgrid, 256, 0, stream, w, (__half*)ws, dgrad, J, K);
    wp = ws;
  }
  R3Params p{};
  p.N = n; p.H = h; p.W = wd; p.K = K; p.J = J;
  p.KS = K % 64 == 0 ? 64 : (K % 32 == 0 ? 32 : 16);
  p.strips = b2u_cdiv(wd, kSW);
  p.y = (__half*)y; p.ldy = ldy; p.bias = bias; p.act = act;
  p.mask = (const __half*)mask; p.ldmask = ldmask; p.mask_act = mask_act; p.accumulate = accumulate;
  p.stats = stats; p.colsum = colsum;
  p.bits_out = (uint8_t*)relu_bits_out;
  p.dbg = g_b2u_dbg;
  if (mask != nullptr && mask_act == B2U_ACT_RELU_BITS) {
    p.bits_in = (const uint8_t*)mask;
    p.mask = nullptr;
    mask = nullptr;
  }
  const uint32_t rowb = p.KS * 2;
  const int kslabs = K / p.KS, NT = 3 * J;
  p.a_stage = (uint32_t)(((size_t)kPatch * rowb + 1023) & ~(size_t)1023);
  p.b_tile = (uint32_t)(((size_t)NT * rowb + 1023) & ~(size_t)1023);
  const size_t wres = 3 * (size_t)kslabs * p.b_tile;
  const size_t tail = (2 * kMaxSAr + 1 + 2 * kNS) * 8 + 32 + (size_t)3 * J * 4 + 64;
  p.SA = (int)((210 * 1024 - 1024 - wres - tail) / p.a_stage);
  if (p.SA > kMaxSAr) p.SA = kMaxSAr;
  B2U_REQUIRE(p.SA >= 2, "tc_conv3r: tiles do not fit shared memory (K=%d J=%d)", K, J);
  size_t smem = 1024 + (size_t)p.SA * p.a_stage + wres + tail;
  if (smem < 120 * 1024) smem = 120 * 1024;      // never two CTAs on an SM: the accumulator ring takes the whole TMEM
  R3Maps maps;
  {
    cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)wd * ldx * 2, (cuuint64_t)h * wd * ldx * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.KS, (cuuint32_t)kPatch, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = p.KS == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (p.KS == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = g_encr(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b2u_set_error("tc_conv3r: activation tensor map failed (%d)", (int)r); return B2U_ERR_CUDA; }
    // the packed bank [t = dh*3 + dw][J][K] viewed as (K, J, dw, dh): a box of one dw takes its three dh taps, i.e. the
    // rows dh*J + co of the N = 3J operand
    cuuint64_t bd[4] = {(cuuint64_t)K, (cuuint64_t)J, 3, 3};
    cuuint64_t bs[3] = {(cuuint64_t)K * 2, (cuuint64_t)K * J * 2, (cuuint64_t)3 * K * J * 2};
    cuuint32_t bb[4] = {(cuuint32_t)p.KS, (cuuint32_t)J, 1, 3};
    r = g_encr(&maps.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(wp), bd, bs, bb, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b2u_set_error("tc_conv3r: weight tensor map failed (%d)", (int)r); return B2U_ERR_CUDA; }
  }
  if (!g_attrr) {
#define B2U_R3_ATTR(F) B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3r_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    B2U_R3_ATTR(0); B2U_R3_ATTR(1); B2U_R3_ATTR(2); B2U_R3_ATTR(3); B2U_R3_ATTR(4); B2U_R3_ATTR(6); B2U_R3_ATTR(8); B2U_R3_ATTR(10);
#undef B2U_R3_ATTR
    g_attrr = true;
  }
  const long long units = (long long)n * p.strips * h;
  B2U_REQUIRE(units < (1LL << 40), "tc_conv3r: too many rows");
  const int grid = (int)(units < B2U_NUM_SMS ? units : B2U_NUM_SMS);
  const int flags = ((mask != nullptr || accumulate) ? kR_MASKACC : 0) | ((stats != nullptr || colsum != nullptr) ? kR_SUMS : 0) |
                    (p.bits_out != nullptr ? kR_BITS_OUT : 0) | (p.bits_in != nullptr ? kR_BITS_IN : 0);
  switch (flags) {
    case 0: B2U_LAUNCH(tc_conv3r_kernel<0>, grid, kThreadsR, smem, stream, maps, p); break;
    case 1: B2U_LAUNCH(tc_conv3r_kernel<1>, grid, kThreadsR, smem, stream, maps, p); break;
    case 2: B2U_LAUNCH(tc_conv3r_kernel<2>, grid, kThreadsR, smem, stream, maps, p); break;
    case 3: B2U_LAUNCH(tc_conv3r_kernel<3>, grid, kThreadsR, smem, stream, maps, p); break;
    case 4: B2U_LAUNCH(tc_conv3r_kernel<4>, grid, kThreadsR, smem, stream, maps, p); break;
    case 6: B2U_LAUNCH(tc_conv3r_kernel<6>, grid, kThreadsR, smem, stream, maps, p); break;
    case 8: B2U_LAUNCH(tc_conv3r_kernel<8>, grid, kThreadsR, smem, stream, maps, p); break;
    case 10: B2U_LAUNCH(tc_conv3r_kernel<10>, grid, kThreadsR, smem, stream, maps, p); break;
    default:
      b2u_set_error("tc_conv3r: unsupported feature combination %d (1-bit masks do not combine with accumulate / fp16 masks)", flags);
      return B2U_ERR_ARG;
  }
  return B2U_OK;
}
