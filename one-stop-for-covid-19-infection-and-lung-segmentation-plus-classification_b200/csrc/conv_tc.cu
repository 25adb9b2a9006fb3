// tcgen05 implicit-GEMM convolution for sm_100a (fp16 operands, fp32 accumulation in TMEM).
//
//   out[p][j] = epilogue( sum_t sum_k  A_t[p + off_t][k] * Wp[t][j][k] )
//
// with p a pixel of an NHWC activation tensor, k the reduction channels, j the output channels and
// t the filter taps.  One kernel serves
//   * Conv2D 3x3 'same' forward        (9 taps, offsets -1..1, A = x,  Wp[t][co][ci] = w[t][ci][co])
//   * its data gradient                (9 taps, offsets -1..1, A = dy, Wp[t][ci][co] = w[8-t][ci][co])
//   * Conv2DTranspose 2x2/s2 forward   (1 tap, A = x, J = 4*Cout, scatter epilogue)
//   * its data gradient                (4 taps = the 2x2 sub-grids of dy, each its own tensor map)
//
// Structure (one CTA per SM, persistent over output tiles of 128 pixels x JT channels):
//   warp 0      TMA producer: per (tap, 64-channel K slab) one 4-D box load of the activations
//               (8x16 pixel patch x KS channels, out-of-image rows/cols zero-filled by TMA = the conv
//               padding) and one 3-D box load of the packed fp16 weights, into a 4-stage smem ring with
//               128B/64B/32B swizzle
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=JT, K=16) per 16 channels, accumulators
//               double-buffered in TMEM (2 x JT columns); tcgen05.commit releases smem stages / publishes
//               the accumulator
//   warps 2-5   epilogue: tcgen05.ld -> bias / ReLU / ELU / activation-derivative mask / accumulate ->
//               fp16 NHWC stores (strided: writes straight into concat-buffer channel slices) and the
//               per-channel sum / sum-of-squares for the following BatchNorm (warp-shuffle transpose
//               reduction -> smem -> one fp64 atomic per channel per CTA)
#include <cuda.h>
#include "common.cuh"
#include "internal.h"
#include "launch.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kMaxStages = 12;
constexpr int kTileH = 8, kTileW = 16, kTileM = 128;       // 8 x 16 output pixels per tile
constexpr int kEpiWarpsTc = 16;                            // four per TMEM lane group (column quarters): the epilogue
                                                           // is instruction-latency bound, more warps hide it
constexpr int kThreadsTc = 64 + 32 * kEpiWarpsTc;
constexpr int kMaxTaps = 9;

struct TcParams {
  int N, H, W;            // output-pixel grid of the GEMM (for convT fwd: the INPUT grid)
  int K, J;               // reduction channels, output columns
  int KS, JT;             // K slab (64/32/16) and N tile
  int stages;             // smem ring depth (as many as fit: small-channel layers are TMA-latency bound)
  int ntaps;
  int tap_dh[kMaxTaps], tap_dw[kMaxTaps], tap_map[kMaxTaps];   // pixel offset and tensor-map index per tap
  int mode;               // 0: plain NHWC store, 1: convT scatter (column = (a*2+b)*Cout + co)
  int cout;               // mode 1: Cout
  __half* y; int ldy;
  const float* bias;      // [J] (mode 1: [Cout]) or null
  int act;
  const __half* mask; int ldmask; int mask_act;
  int accumulate;
  double* stats;          // sums at stats[c], squares at stats[stats_sq_off + c]; null: none
  int stats_sq_off;
  float* colsum;          // per-column sums of the stored values (fp32 atomics): the bias gradient of the layer whose
                          // output gradient this kernel writes; null: none
};

struct TcMaps {
  CUtensorMap a[4];
  CUtensorMap b;
};

__device__ __forceinline__ float apply_act(float v, int act) { return act_fwd(v, act); }

// lane j ends with sum over lanes of v[j] (32 values per lane), 31 shuffles
__device__ __forceinline__ float transpose_reduce16(float v[16], int lane) {
  // 16 columns x 32 lanes: first fold lanes 16 apart (values stay 16 wide), then transpose-reduce
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
  return v[0];      // lane (l & 15) owns column (l & 15); both half-warps hold the same totals
}

// kRegStats: BatchNorm statistics accumulate in registers over ALL tiles of the persistent CTA (a thread owns
// <= 32 fixed columns; requires gridDim.x % (J / JT) == 0 so that the CTA always sees the same N tile) and are
// reduced across lanes once at the end, instead of two 31-shuffle transposes per 16 columns per tile.
// kRegStats = number of columns a thread keeps statistics for (0: none / shared-memory path, 16, 32): with 576 threads the
// kernel has 96 registers per thread, and 2 x 32 statistics registers beside the 16-column chunk spilled ~30 of them
// (ncu: 11 % of the stall samples on LDL / STL, profiles/r2_convt_*_ncu_source_top.txt) -- the host picks N tiles that
// leave 16 columns per thread wherever it can.
template <int kRegStats>
__global__ void __launch_bounds__(kThreadsTc, 1) tc_conv_kernel(const __grid_constant__ TcMaps maps,
                                                                 const __grid_constant__ TcParams prm) {
  B2U_PDL_LAUNCH_DEPENDENTS();      // B2U_PDL_WAIT() follows the CTA-local setup (barriers, TMEM)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A tile 128 x KS halves][B tile JT x KS halves] | barriers | tmem ptr | bias | stats
  const int KS = prm.KS, JT = prm.JT;
  const uint32_t a_bytes = kTileM * KS * 2, b_bytes = JT * KS * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;          // both multiples of 1024 (KS>=16, JT>=16 -> b>=512!)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t b_stride = (b_bytes + 1023) & ~1023u;
  const uint32_t stage_stride = a_bytes + b_stride;
  const int kStages = prm.stages;
  uint8_t* tail = smem + kStages * stage_stride;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 4) + 15) & ~(uintptr_t)15);
  float* s_stats = s_bias + prm.J;                         // [2*J]
  (void)stage_bytes;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (role dispatch)
  const int lane = threadIdx.x & 31;
  const int tiles_w = (prm.W + kTileW - 1) / kTileW, tiles_h = (prm.H + kTileH - 1) / kTileH;
  const int nj = prm.J / JT;
  const unsigned ntiles = (unsigned)(prm.N * tiles_h * tiles_w * nj);   // host guarantees < 2^31
  const int kslabs = prm.K / KS;
  const int ksteps = prm.ntaps * kslabs;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * JT)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&tfull_bar[s], 1); tc::mbar_init(&tempty_bar[s], kEpiWarpsTc); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, tmem_cols);
  B2U_PDL_WAIT();                    // everything below may read what the preceding kernel wrote
  for (int i = threadIdx.x; i < prm.J; i += kThreadsTc) {
    float b = 0.f;
    if (prm.bias != nullptr) b = prm.bias[prm.mode == 1 ? (i % prm.cout) : i];
    s_bias[i] = b;
  }
  for (int i = threadIdx.x; i < 2 * prm.J; i += kThreadsTc) s_stats[i] = 0.f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer =========================================
    if (lane == 0) {
      tc::prefetch_tmap(&maps.b);
      tc::prefetch_tmap(&maps.a[0]);
      int stage = 0;
      uint32_t phase = 0;
      for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int jt = (int)(tile % (unsigned)nj);
        const unsigned pt = tile / (unsigned)nj;
        const int tw = (int)(pt % (unsigned)tiles_w);
        const unsigned r = pt / (unsigned)tiles_w;
        const int th = (int)(r % (unsigned)tiles_h);
        const int n = (int)(r / (unsigned)tiles_h);
        for (int t = 0; t < prm.ntaps; ++t) {
          const CUtensorMap* am = &maps.a[prm.tap_map[t]];
          for (int ks = 0; ks < kslabs; ++ks) {
            tc::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * stage_stride;
            uint8_t* sb = sa + a_bytes;
            tc::mbar_expect_tx(&full_bar[stage], a_bytes + b_bytes);
            tc::tma_load_4d(sa, am, &full_bar[stage], ks * KS, tw * kTileW + prm.tap_dw[t], th * kTileH + prm.tap_dh[t], n);
            tc::tma_load_3d(sb, &maps.b, &full_bar[stage], ks * KS, jt * JT, t);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    {   // whole warp in uniform control flow, one elected lane issues (tc_common.cuh)
      const uint32_t idesc = tc::idesc_f16(kTileM, JT, 0, 0);
      const uint64_t layout = KS == 64 ? tc::SWZ_128B : (KS == 32 ? tc::SWZ_64B : tc::SWZ_32B);
      const uint32_t sbo = 8 * KS * 2;                     // 8 rows of KS halves
      // descriptor words hoisted out of the issue loop (single-thread dependent instruction stream)
      const uint32_t ab_hi = (uint32_t)(tc::smem_desc(0, 16, sbo, layout) >> 32);
      const uint32_t ab_lo0 = ((tc::smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t stage16 = stage_stride >> 4, a_bytes16 = a_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * JT;
        for (int k = 0; k < ksteps; ++k) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::fence_after_sync();
          const uint32_t a_lo = ab_lo0 + (uint32_t)stage * stage16;
          const uint32_t b_lo = a_lo + a_bytes16;
          for (int kk = 0; kk < KS / 16; ++kk) {
            const uint64_t ad = ((uint64_t)ab_hi << 32) | (uint64_t)(a_lo + 2 * kk);
            const uint64_t bd = ((uint64_t)ab_hi << 32) | (uint64_t)(b_lo + 2 * kk);
            tc::mma_f16_ss_elect(d_tmem, ad, bd, idesc, (k | kk) != 0);
          }
          tc::mma_commit_elect(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc::mma_commit_elect(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue ==============================================
    // 16 warps: TMEM lane group = warp & 3, column part = (warp - 2) >> 2 of `nparts` equal parts of the N tile
    const int lg = warp & 3;
    const int part = (warp - 2) >> 2;
    const int row = lg * 32 + lane;                        // tile row = pixel (row / 16, row % 16)
    const int nparts = JT % 64 == 0 ? 4 : (JT % 32 == 0 ? 2 : 1);
    const int ccols = JT / nparts;
    const int cbeg = part * ccols;
    const bool has_cols = part < nparts;
    float rs1[kRegStats ? kRegStats : 1], rs2[kRegStats ? kRegStats : 1];
#pragma unroll
    for (int i = 0; i < (kRegStats ? kRegStats : 1); ++i) { rs1[i] = 0.f; rs2[i] = 0.f; }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int jt = (int)(tile % (unsigned)nj);
      const unsigned pt = tile / (unsigned)nj;
      const int tw = (int)(pt % (unsigned)tiles_w);
      const unsigned r = pt / (unsigned)tiles_w;
      const int th = (int)(r % (unsigned)tiles_h);
      const int n = (int)(r / (unsigned)tiles_h);
      const int h = th * kTileH + row / kTileW, w = tw * kTileW + row % kTileW;
      const bool valid = h < prm.H && w < prm.W;
      const long long pix = ((long long)n * prm.H + h) * prm.W + w;
      // activation-derivative mask of this thread's (<= 32) columns: loaded BEFORE the accumulator wait, so that its DRAM
      // round trip overlaps the wait instead of following it chunk by chunk (ncu: 26 % of the stall samples sat on the
      // first use of the mask in the transposed-conv data gradient)
      // (the 32-column statistics variant has no registers to spare for it; the 16-column one holds 16 columns of mask)
      constexpr int kMk = kRegStats == 16 ? 2 : 4;
      const bool mask_early = kRegStats != 32 && prm.mask != nullptr && ccols <= 8 * kMk && has_cols && valid;
      uint4 mk[kMk];
      if (mask_early) {
        const uint4* mp = reinterpret_cast<const uint4*>(prm.mask + pix * prm.ldmask + jt * JT + cbeg);
        mk[0] = __ldg(mp);
        mk[1] = __ldg(mp + 1);
        if constexpr (kMk == 4) {
          if (ccols > 16) { mk[2] = __ldg(mp + 2); mk[3] = __ldg(mp + 3); }
        }
      }
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::fence_after_sync();
      if (has_cols) {
#pragma unroll 2
        for (int cc = 0; cc < ccols; cc += 16) {
          const int c0 = cbeg + cc;
          float v[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * JT + c0, v);
          const int j0 = jt * JT + c0;
          const float4* bp = reinterpret_cast<const float4*>(s_bias + j0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = bp[q];
            v[4 * q + 0] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
          }
          if (prm.act == B2U_ACT_RELU) {          // activation switch outside the element loop (branch-free chains)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (prm.act == B2U_ACT_ELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
          }
          __half* dst;
          if (prm.mode == 1) {
            int ab = j0 / prm.cout, co = j0 - ab * prm.cout;
            long long op = (((long long)n * 2 * prm.H + 2 * h + (ab >> 1)) * (2LL * prm.W) + 2 * w + (ab & 1));
            dst = prm.y + op * prm.ldy + co;
          } else {
            dst = prm.y + pix * prm.ldy + j0;
          }
          if (valid) {
            if (mask_early) {
              float m[8];
              unpack8h(cc == 0 ? mk[0] : mk[kMk == 4 ? 2 : 0], m);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= act_bwd_from_y(m[i], prm.mask_act);
              unpack8h(cc == 0 ? mk[1] : mk[kMk == 4 ? 3 : 1], m);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 + i] *= act_bwd_from_y(m[i], prm.mask_act);
            } else if (prm.mask != nullptr) {
              float m[8];
              load8<__half>(prm.mask + pix * prm.ldmask + j0, m);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= act_bwd_from_y(m[i], prm.mask_act);
              load8<__half>(prm.mask + pix * prm.ldmask + j0 + 8, m);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 + i] *= act_bwd_from_y(m[i], prm.mask_act);
            }
            if (prm.accumulate) {
              float e[8];
              load8<__half>(dst, e);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += e[i];
              load8<__half>(dst + 8, e);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 + i] += e[i];
            }
            store8<__half>(dst, v);
            store8<__half>(dst + 8, v + 8);
          }
          if (kRegStats) {                       // (launched only when statistics or column sums are wanted)
            if (valid) {
              if (kRegStats < 32 || cc == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { rs1[i] += v[i]; rs2[i] = fmaf(v[i], v[i], rs2[i]); }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) { rs1[(kRegStats == 32 ? 16 : 0) + i] += v[i]; rs2[(kRegStats == 32 ? 16 : 0) + i] = fmaf(v[i], v[i], rs2[(kRegStats == 32 ? 16 : 0) + i]); }
              }
            }
          } else if (prm.stats != nullptr || prm.colsum != nullptr) {
            const int sc = prm.mode == 1 ? (j0 % prm.cout) : j0;         // scatter mode: column -> output channel
            float q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i] = valid ? v[i] : 0.f;
            if (prm.stats != nullptr) {
              float sq[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) sq[i] = q[i] * q[i];
              float s2 = transpose_reduce16(sq, lane);
              if (lane < 16) atomicAdd(&s_stats[prm.J + sc + lane], s2);
            }
            float s1 = transpose_reduce16(q, lane);
            if (lane < 16) atomicAdd(&s_stats[sc + lane], s1);
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (kRegStats && has_cols && blockIdx.x < ntiles) {
      // the CTA's N tile is fixed (gridDim.x % nj == 0): registers hold columns [cbeg, cbeg + ccols) of it
      const int jt = (int)(blockIdx.x % (unsigned)nj);
      for (int cc = 0; cc < ccols; cc += 16) {
        float q[16], sq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          q[i] = cc == 0 ? rs1[i] : rs1[(kRegStats == 32 ? 16 : 0) + i];
          sq[i] = cc == 0 ? rs2[i] : rs2[(kRegStats == 32 ? 16 : 0) + i];
        }
        float s1 = transpose_reduce16(q, lane);
        float s2 = transpose_reduce16(sq, lane);
        if (lane < 16) {
          const int j0 = jt * JT + cbeg + cc;
          const int sc = prm.mode == 1 ? (j0 % prm.cout) : j0;
          atomicAdd(&s_stats[sc + lane], s1);
          atomicAdd(&s_stats[prm.J + sc + lane], s2);
        }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (prm.stats != nullptr) {
    const int nstat = prm.mode == 1 ? prm.cout : prm.J;
    for (int i = threadIdx.x; i < nstat; i += kThreadsTc) {
      atomicAdd(&prm.stats[i], (double)s_stats[i]);
      atomicAdd(&prm.stats[prm.stats_sq_off + i], (double)s_stats[prm.J + i]);
    }
  }
  if (prm.colsum != nullptr) {
    const int nstat = prm.mode == 1 ? prm.cout : prm.J;
    for (int i = threadIdx.x; i < nstat; i += kThreadsTc) {
      const float sv = s_stats[i];
      if (sv != 0.f) atomicAdd(&prm.colsum[i], sv);
    }
  }
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// ---- weight packing: fp32 Keras layouts -> fp16 [tap][j][k] ---------------------------------------
// mode 0 (conv fwd)   : Wp[t][co][ci] = w[t][ci][co]
// mode 1 (conv dgrad) : Wp[t][ci][co] = w[8-t][ci][co]
// mode 2 (convT fwd)  : Wp[0][q][ci]  = w[q][ci]                 (q = (a*2+b)*Cout + co)
// mode 3 (convT dgrad): Wp[ab][ci][co] = w[ab][co][ci]
__global__ void pack_weights_kernel(const float* __restrict__ w, __half* __restrict__ wp, int mode, int taps, int J,
                                    int K) {
  B2U_PDL_PROLOGUE();
  long long total = (long long)taps * J * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    long long r = i / K;
    int j = (int)(r % J);
    int t = (int)(r / J);
    float v;
    if (mode == 0) v = w[((long long)t * K + k) * J + j];
    else if (mode == 1) v = w[((long long)(8 - t) * J + j) * K + k];
    else if (mode == 2) v = w[(long long)j * K + k];
    else v = w[((long long)t * K + k) * J + j];
    wp[i] = __float2half_rn(v);
  }
}

// all layers of a model in one launch (b2u_pack_weights): table of 8 x int64 records, see include/b200unet.h.
// Work unit = one 32 x 32 tile of one tap matrix, staged through shared memory so that both the fp32 reads and the
// fp16 writes are coalesced whether the mode transposes (0, 3) or copies (1, 2).
__global__ void __launch_bounds__(256) pack_all_kernel(const long long* __restrict__ tab, int n_entries,
                                                       const float* __restrict__ params, __half* __restrict__ wpack,
                                                       long long total_tiles) {
  B2U_PDL_PROLOGUE();
  __shared__ long long st[128 * 8];
  __shared__ float tile[32][33];
  for (int i = threadIdx.x; i < n_entries * 8; i += blockDim.x) st[i] = tab[i];
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (long long u = blockIdx.x; u < total_tiles; u += gridDim.x) {
    int lo = 0, hi = n_entries - 1;                                 // last entry whose first tile is <= u
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (st[mid * 8 + 2] <= u) lo = mid; else hi = mid - 1;
    }
    const long long* e = st + lo * 8;
    const float* w = params + e[0];
    __half* dst = wpack + e[1];
    const int mode = (int)e[3], J = (int)e[5], K = (int)e[6];
    const int Ks = e[7] > 0 ? (int)e[7] : K;                        // source K (< K: the packed copy is zero-padded in K)
    const int tj = (J + 31) >> 5, tk = (K + 31) >> 5;
    const int lu = (int)(u - e[2]);
    const int t = lu / (tj * tk), r = lu % (tj * tk);
    const int j0 = (r / tk) * 32, k0 = (r % tk) * 32;
    const bool transpose = mode == 0 || mode == 3;                  // source tap matrix is [K][J] (else [J][K])
    const float* src = w + (long long)(mode == 1 ? 8 - t : t) * J * Ks;
    __syncthreads();                                                // previous tile's readers are done
    if (transpose) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {                                 // rows = k, columns = j (coalesced in j)
        const int k = k0 + ty + 8 * i, j = j0 + tx;
        tile[ty + 8 * i][tx] = (k < Ks && j < J) ? src[(long long)k * J + j] : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {                                 // rows = j, columns = k (coalesced in k)
        const int j = j0 + ty + 8 * i, k = k0 + tx;
        tile[tx][ty + 8 * i] = (j < J && k < Ks) ? src[(long long)j * Ks + k] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {                                   // dst[t][j][k], coalesced in k; tile[k][j]
      const int j = j0 + ty + 8 * i, k = k0 + tx;
      if (j < J && k < K) dst[((long long)t * J + j) * K + k] = __float2half_rn(tile[tx][ty + 8 * i]);
    }
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int get_encode() {
  if (g_encode != nullptr) return B2U_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B2U_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    b2u_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B2U_ERR_CUDA;
  }
  g_encode = (EncodeTiledFn)fn;
  return B2U_OK;
}

CUtensorMapSwizzle swizzle_for(int ks) {
  return ks == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (ks == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// activation map: dims (C, W, H, N), element strides given in elements of the underlying fp16 buffer
int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sW, long long sH,
                 long long sN, int KS) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)KS, (cuuint32_t)kTileW, (cuuint32_t)kTileH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(KS), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b2u_set_error("cuTensorMapEncodeTiled(activations C=%d W=%d H=%d N=%d sW=%lld KS=%d) failed: %d", C, W, H, N, sW, KS,
                  (int)r);
    return B2U_ERR_CUDA;
  }
  return B2U_OK;
}

int make_w_map(CUtensorMap* m, const void* base, int K, int J, int taps, int KS, int JT) {
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)J, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * J * 2};
  cuuint32_t box[3] = {(cuuint32_t)KS, (cuuint32_t)JT, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(KS), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b2u_set_error("cuTensorMapEncodeTiled(weights K=%d J=%d taps=%d KS=%d JT=%d) failed: %d", K, J, taps, KS, JT, (int)r);
    return B2U_ERR_CUDA;
  }
  return B2U_OK;
}

int pick_ks(int K) { return K % 64 == 0 ? 64 : (K % 32 == 0 ? 32 : (K % 16 == 0 ? 16 : 0)); }
int pick_jt(int J) {
  if (J <= 256) return J;
  for (int jt = 256; jt >= 16; jt -= 16)
    if (J % jt == 0) return jt;
  return 0;
}

size_t tail_bytes_for(int J) { return (2 * kMaxStages + 4) * 8 + 32 + (size_t)3 * J * 4 + 64; }
int stages_for(int KS, int JT, int J) {
  size_t a = (size_t)kTileM * KS * 2, b = ((size_t)JT * KS * 2 + 1023) & ~(size_t)1023;
  long long room = 220 * 1024 - 1024 - (long long)tail_bytes_for(J);
  int st = (int)(room / (long long)(a + b));
  return st > kMaxStages ? kMaxStages : st;
}
size_t smem_bytes_for(int KS, int JT, int J, int stages) {
  size_t a = (size_t)kTileM * KS * 2, b = ((size_t)JT * KS * 2 + 1023) & ~(size_t)1023;
  return 1024 + stages * (a + b) + tail_bytes_for(J);
}

bool g_attr_set = false;

int launch_tc(const TcMaps& maps, TcParams& prm, void* stream) {
  prm.stages = stages_for(prm.KS, prm.JT, prm.J);
  B2U_REQUIRE(prm.stages >= 2, "tc_conv: tile too large for a 2-stage pipeline");
  size_t smem = smem_bytes_for(prm.KS, prm.JT, prm.J, prm.stages);
  B2U_REQUIRE(smem <= 227 * 1024, "tc_conv: shared memory %zu exceeds 227 KB", smem);
  if (!g_attr_set) {
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_attr_set = true;
  }
  const int nj = prm.J / prm.JT;
  long long tiles = (long long)prm.N * b2u_cdiv(prm.H, kTileH) * b2u_cdiv(prm.W, kTileW) * nj;
  B2U_REQUIRE(tiles < (1LL << 31), "tc_conv: too many tiles");
  int grid = (int)(tiles < B2U_NUM_SMS ? tiles : B2U_NUM_SMS);
  // register statistics need a fixed N tile per CTA (grid a multiple of nj) and <= 32 columns per epilogue thread
  const int nparts = prm.JT % 64 == 0 ? 4 : (prm.JT % 32 == 0 ? 2 : 1);
  bool reg_stats = (prm.stats != nullptr || prm.colsum != nullptr) && prm.JT / nparts <= 32 && nj <= B2U_NUM_SMS;
  if (reg_stats) {
    const int g2 = grid / nj * nj;
    if (g2 >= 1) grid = g2; else reg_stats = false;
  }
  if (reg_stats && prm.JT / nparts <= 16) { B2U_LAUNCH(tc_conv_kernel<16>, grid, kThreadsTc, smem, stream, maps, prm); }
  else if (reg_stats) { B2U_LAUNCH(tc_conv_kernel<32>, grid, kThreadsTc, smem, stream, maps, prm); }
  else { B2U_LAUNCH(tc_conv_kernel<0>, grid, kThreadsTc, smem, stream, maps, prm); }
  return B2U_OK;
}

int pack(const float* w, void* ws, size_t ws_bytes, int mode, int taps, int J, int K, void* stream) {
  size_t need = (size_t)taps * J * K * 2;
  B2U_REQUIRE(ws != nullptr && need <= ws_bytes, "tc_conv: workspace too small (%zu > %zu)", need, ws_bytes);
  long long total = (long long)taps * J * K;
  int grid = (int)((total + 255) / 256);
  if (grid > 8 * B2U_NUM_SMS) grid = 8 * B2U_NUM_SMS;
  B2U_LAUNCH(pack_weights_kernel, grid, 256, 0, stream, w, (__half*)ws, mode, taps, J, K);
  return B2U_OK;
}

}  // namespace

int b2u_tc_compiled(void) { return 1; }

extern "C" int b2u_pack_weights(const long long* d_table, int n_entries, const float* params, void* wpack,
                                long long total, void* stream) {
  B2U_REQUIRE(d_table != nullptr && params != nullptr && wpack != nullptr && n_entries >= 1 && n_entries <= 128 && total > 0,
              "pack_weights: bad arguments (1..128 table entries)");
  long long grid = total < 8LL * B2U_NUM_SMS ? total : 8LL * B2U_NUM_SMS;
  B2U_LAUNCH(pack_all_kernel, (int)grid, 256, 0, stream, d_table, n_entries, params, (__half*)wpack, total);
  return B2U_OK;
}

int b2u_tc_conv3x3_ok(int k, int j, int ld_in, int ld_out) {
  return pick_ks(k) != 0 && j % 16 == 0 && pick_jt(j) != 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && j <= 1024;
}
int b2u_tc_convt_ok(int cin, int cout, int ld_small, int ld_big) {
  return pick_ks(cin) != 0 && pick_ks(cout) != 0 && cout % 16 == 0 && cin % 16 == 0 && pick_jt(4 * cout) != 0 &&
         ld_small % 8 == 0 && ld_big % 8 == 0 && 4 * cout <= 1024 && cin <= 1024;
}

int b2u_tc_conv3x3(const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act, void* y,
                   int ldy, int J, double* stats, float* colsum, const void* mask, int ldmask, int mask_act,
                   int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp, void* stream,
                   void* relu_bits_out) {
  B2U_REQUIRE(relu_bits_out == nullptr && mask_act != B2U_ACT_RELU_BITS, "tc_conv3x3 (per-tap): 1-bit masks need the halo kernel");
  int rc = get_encode();
  if (rc != B2U_OK) return rc;
  TcParams p{};
  p.N = n; p.H = h; p.W = wd; p.K = K; p.J = J; p.KS = pick_ks(K); p.JT = pick_jt(J); p.ntaps = 9;
  for (int t = 0; t < 9; ++t) { p.tap_dh[t] = t / 3 - 1; p.tap_dw[t] = t % 3 - 1; p.tap_map[t] = 0; }
  p.mode = 0; p.cout = J; p.y = (__half*)y; p.ldy = ldy; p.bias = bias; p.act = act;
  p.mask = (const __half*)mask; p.ldmask = ldmask; p.mask_act = mask_act; p.accumulate = accumulate; p.stats = stats;
  p.stats_sq_off = J; p.colsum = colsum;
  if (colsum != nullptr && J % 128 == 0 && p.JT > 128) p.JT = 128;
  if (wp == nullptr) {
    rc = pack(w, ws, ws_bytes, dgrad ? 1 : 0, 9, J, K, stream);
    if (rc != B2U_OK) return rc;
    wp = ws;
  }
  TcMaps maps;
  rc = make_act_map(&maps.a[0], x, K, wd, h, n, ldx, (long long)wd * ldx, (long long)h * wd * ldx, p.KS);
  if (rc != B2U_OK) return rc;
  for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
  rc = make_w_map(&maps.b, wp, K, J, 9, p.KS, p.JT);
  if (rc != B2U_OK) return rc;
  return launch_tc(maps, p, stream);
}

// N tile of the transposed-conv kernels when their epilogue keeps statistics / reads a mask: 64 (16 columns per thread, no
// register spills) measured no faster than 128 on B200 (forward 0.093 vs 0.091 ms at 256^2 x 64 -> 32; data gradient of the
// 128-channel level 0.054 vs 0.049): the spills were not what bounds these epilogues.  Default 128.
int g_b2u_convt_jt = 128;

int b2u_tc_convt_fwd(const void* x, int ldx, int cin, const float* w, const float* bias, void* y, int ldy, int cout,
                     double* stats, int stats_sq_off, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp,
                     void* stream) {
  int rc = get_encode();
  if (rc != B2U_OK) return rc;
  TcParams p{};
  p.N = n; p.H = h; p.W = wd; p.K = cin; p.J = 4 * cout; p.KS = pick_ks(cin); p.JT = pick_jt(4 * cout);
  // statistics in registers: 64-column N tiles leave 16 columns per epilogue thread (no spills, see tc_conv_kernel);
  // `convt_jt` = 128 restores round 1's 32 columns per thread
  if (stats != nullptr && p.J % 128 == 0) p.JT = (g_b2u_convt_jt == 64 && p.J % 64 == 0) ? 64 : 128;
  if (p.JT > cout && p.JT % cout != 0) p.JT = cout;       // a 16-column chunk must not straddle two (a,b) groups
  p.ntaps = 1; p.tap_dh[0] = 0; p.tap_dw[0] = 0; p.tap_map[0] = 0;
  p.mode = 1; p.cout = cout; p.y = (__half*)y; p.ldy = ldy; p.bias = bias; p.act = B2U_ACT_NONE;
  p.stats = stats; p.stats_sq_off = stats_sq_off;
  if (wp == nullptr) {
    rc = pack(w, ws, ws_bytes, 2, 1, 4 * cout, cin, stream);
    if (rc != B2U_OK) return rc;
    wp = ws;
  }
  TcMaps maps;
  rc = make_act_map(&maps.a[0], x, cin, wd, h, n, ldx, (long long)wd * ldx, (long long)h * wd * ldx, p.KS);
  if (rc != B2U_OK) return rc;
  for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
  rc = make_w_map(&maps.b, wp, cin, 4 * cout, 1, p.KS, p.JT);
  if (rc != B2U_OK) return rc;
  return launch_tc(maps, p, stream);
}

int b2u_tc_convt_dgrad(const void* dy, int lddy, int cout, const float* w, void* dx, int lddx, int cin,
                       const void* mask, int ldmask, int mask_act, int accumulate, float* colsum, int n, int h, int wd,
                       void* ws, size_t ws_bytes, const void* wp, void* stream) {
  int rc = get_encode();
  if (rc != B2U_OK) return rc;
  TcParams p{};
  p.N = n; p.H = h; p.W = wd; p.K = cout; p.J = cin; p.KS = pick_ks(cout); p.JT = pick_jt(cin); p.ntaps = 4;
  for (int t = 0; t < 4; ++t) { p.tap_dh[t] = 0; p.tap_dw[t] = 0; p.tap_map[t] = t; }
  p.mode = 0; p.cout = cin; p.y = (__half*)dx; p.ldy = lddx; p.bias = nullptr; p.act = B2U_ACT_NONE;
  p.mask = (const __half*)mask; p.ldmask = ldmask; p.mask_act = mask_act; p.accumulate = accumulate;
  p.colsum = colsum;
  if (colsum != nullptr && cin % 128 == 0 && p.JT > 128) p.JT = 128;   // <= 32 columns per epilogue thread
  if ((colsum != nullptr || mask != nullptr) && g_b2u_convt_jt == 64 && cin % 64 == 0 && p.JT > 64) p.JT = 64;   // 16 per thread
  if (wp == nullptr) {
    rc = pack(w, ws, ws_bytes, 3, 4, cin, cout, stream);
    if (rc != B2U_OK) return rc;
    wp = ws;
  }
  TcMaps maps;
  // tap (a,b): the sub-grid dy[n, 2i+a, 2j+b, :] seen as an (N,H,W,C) tensor with doubled pixel strides
  for (int t = 0; t < 4; ++t) {
    const __half* base = (const __half*)dy + ((long long)(t >> 1) * (2 * wd) + (t & 1)) * lddy;
    rc = make_act_map(&maps.a[t], base, cout, wd, h, n, 2LL * lddy, 4LL * wd * lddy, 4LL * h * wd * lddy, p.KS);
    if (rc != B2U_OK) return rc;
  }
  rc = make_w_map(&maps.b, wp, cout, cin, 4, p.KS, p.JT);
  if (rc != B2U_OK) return rc;
  return launch_tc(maps, p, stream);
}
