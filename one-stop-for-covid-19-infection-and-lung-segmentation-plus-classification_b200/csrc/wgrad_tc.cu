// tcgen05 weight-gradient kernel for sm_100a: dW = A^T * B reduced over pixels.
//
//   D[r][j] = sum_p  A_r[p] * B[p][j]        M = 128 accumulator rows, N = JT columns, K = pixels
//
// Both operands are NHWC activation tiles ([128 pixels][channels], channels contiguous), i.e. "MN-major"
// for the tensor core: the pixel axis is the GEMM K axis, so the same TMA boxes the forward kernel uses
// feed tcgen05.mma with a_major = b_major = MN.  The 128 accumulator rows are assembled from `nchunks`
// activation tiles of KSA channels each (KSA * nchunks = 128): for thin layers several filter taps share
// one MMA (4 taps x 32 channels, 2 taps x 64 channels), for wide layers a chunk is a 64-channel slab.
//
//   Conv2D 3x3      : A = x shifted by the tap offset (TMA zero-fill = padding), B = dy;  D -> dw[t][ci][co]
//   Conv2DTranspose : A = the (a,b) sub-grid of dy (own tensor map),            B = x;   D -> dw[a][b][co][ci]
//
// One CTA per SM, persistent over tasks (row group, N tile, pixel split); split-K partial sums are
// added to the fp32 gradient buffer with atomics.  Warp roles as in conv_tc.cu.
#include <cuda.h>
#include "common.cuh"
#include "internal.h"
#include "launch.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kThreadsW = 192;
constexpr int kTileH = 8, kTileW = 16;        // 128 pixels per K block
constexpr int kMaxGroups = 36, kMaxChunks = 8;

struct WgParams {
  int N, H, W;                 // pixel grid both operands are sampled on
  int KSA, nchunks;            // A chunk width (channels) ; 128 / KSA
  int KSB, JT, J;              // B slab width, N tile, total B channels
  int ngroups;                 // row groups (each: nchunks chunks = 128 accumulator rows)
  int8_t ch_map[kMaxGroups][kMaxChunks];     // tensor-map index of the chunk (-1: unused rows)
  int8_t ch_dh[kMaxGroups][kMaxChunks], ch_dw[kMaxGroups][kMaxChunks];
  int16_t ch_c0[kMaxGroups][kMaxChunks];     // first channel of the chunk inside its tensor
  int out_base[kMaxGroups][kMaxChunks];      // dw element offset of the chunk's first row
  int out_row_stride;          // elements between consecutive accumulator rows in dw
  int nsplit;                  // pixel-block splits
  int stages;
  float* dw;
};

struct WgMaps {
  CUtensorMap a[4];
  CUtensorMap b;
};

__global__ void __launch_bounds__(kThreadsW, 1) tc_wgrad_kernel(const __grid_constant__ WgMaps maps,
                                                                 const __grid_constant__ WgParams prm) {
  B2U_PDL_LAUNCH_DEPENDENTS();      // B2U_PDL_WAIT() follows the CTA-local setup (barriers, TMEM)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KSA = prm.KSA, KSB = prm.KSB, JT = prm.JT, nchunks = prm.nchunks, stages = prm.stages;
  const uint32_t a_tile = 128u * KSA * 2, b_tile = 128u * KSB * 2;
  const int nb = JT / KSB;                                   // B slabs per N tile
  const uint32_t a_bytes = a_tile * nchunks, b_bytes = b_tile * nb;   // 32 KB, JT * 256 B
  const uint32_t stage_stride = a_bytes + b_bytes;           // multiples of 1024 (tiles are >= 4 KB)
  uint8_t* tail = smem + stages * stage_stride;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (role dispatch)
  const int lane = threadIdx.x & 31;
  const int tiles_w = (prm.W + kTileW - 1) / kTileW, tiles_h = (prm.H + kTileH - 1) / kTileH;
  const int nblocks = prm.N * tiles_h * tiles_w;             // K blocks of 128 pixels
  const int nj = prm.J / JT;
  const int ntasks = prm.nsplit * prm.ngroups * nj;          // task = (split, group, jt), split-major
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * JT)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&tfull_bar[s], 1); tc::mbar_init(&tempty_bar[s], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, tmem_cols);
  B2U_PDL_WAIT();                    // everything below may read what the preceding kernel wrote
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer =========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int jt = task % nj, g = (task / nj) % prm.ngroups, sp = task / (nj * prm.ngroups);
        for (int blk = sp; blk < nblocks; blk += prm.nsplit) {
          const int tw = blk % tiles_w, th = (blk / tiles_w) % tiles_h, n = blk / (tiles_w * tiles_h);
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_stride;
          uint8_t* sb = sa + a_bytes;
          int live = 0;
          for (int c = 0; c < nchunks; ++c) live += prm.ch_map[g][c] >= 0;
          tc::mbar_expect_tx(&full_bar[stage], live * a_tile + b_bytes);
          for (int c = 0; c < nchunks; ++c) {
            const int m = prm.ch_map[g][c];
            if (m < 0) continue;
            tc::tma_load_4d(sa + c * a_tile, &maps.a[m], &full_bar[stage], prm.ch_c0[g][c],
                            tw * kTileW + prm.ch_dw[g][c], th * kTileH + prm.ch_dh[g][c], n);
          }
          for (int s = 0; s < nb; ++s)
            tc::tma_load_4d(sb + s * b_tile, &maps.b, &full_bar[stage], jt * JT + s * KSB, tw * kTileW, th * kTileH, n);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    {   // whole warp in uniform control flow, one elected lane issues (tc_common.cuh)
      const uint32_t idesc = tc::idesc_f16(128, JT, 1, 1);              // both operands MN-major
      const uint64_t la = KSA == 64 ? tc::SWZ_128B : (KSA == 32 ? tc::SWZ_64B : tc::SWZ_32B);
      const uint64_t lb = KSB == 64 ? tc::SWZ_128B : (KSB == 32 ? tc::SWZ_64B : tc::SWZ_32B);
      const uint32_t rowa = KSA * 2, rowb = KSB * 2;                    // bytes per pixel row of a tile
      const uint32_t a_hi = (uint32_t)(tc::smem_desc(0, a_tile, 8 * rowa, la) >> 32);
      const uint32_t b_hi = (uint32_t)(tc::smem_desc(0, b_tile, 8 * rowb, lb) >> 32);
      const uint32_t a_lbo16 = ((a_tile >> 4) & 0x3FFF) << 16, b_lbo16 = ((b_tile >> 4) & 0x3FFF) << 16;
      const uint32_t a_lo0 = ((tc::smem_u32(smem) & 0x3FFFF) >> 4) | a_lbo16;
      const uint32_t stage16 = stage_stride >> 4, a_bytes16 = a_bytes >> 4;
      const uint32_t ka16 = (16 * rowa) >> 4, kb16 = (16 * rowb) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int sp = task / (nj * prm.ngroups);
        tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * JT;
        uint32_t first = 1;
        for (int blk = sp; blk < nblocks; blk += prm.nsplit) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::fence_after_sync();
          // MN-major descriptors: LBO = distance between channel chunks (tiles), SBO = 8 pixel rows; only the
          // start-address field changes per MMA (16 pixel rows further), so hi words / LBO are hoisted
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage16;
          const uint32_t b_lo = a_lo + a_bytes16;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {                              // 8 x 16 pixels
            const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + kk * ka16);
            const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo - a_lbo16 + b_lbo16 + kk * kb16);
            tc::mma_f16_ss_elect(d_tmem, ad, bd, idesc, (kk != 0) ? 1u : (first ? 0u : 1u));
          }
          first = 0;
          tc::mma_commit_elect(&empty_bar[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        tc::mma_commit_elect(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue ==============================================
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
      const int jt = task % nj, g = (task / nj) % prm.ngroups, sp = task / (nj * prm.ngroups);
      const int c = row / KSA, rr = row % KSA;
      const bool live = prm.ch_map[g][c] >= 0 && sp < nblocks;
      float* out = prm.dw + (long long)prm.out_base[g][c] + (long long)rr * prm.out_row_stride + jt * JT;
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::fence_after_sync();
      for (int c0 = 0; c0 < JT; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * JT + c0, v);
        if (live) {
// 16-byte vector reductions (red.global.add.v4.f32): 4 L2 atomic operations instead of 16
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            atomicAdd(reinterpret_cast<float4*>(out + c0 + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// =====================================================================================================
// HALO variant for thin, full-resolution layers (Cin = 16 / 32 / 64): the layers whose weight gradient is
// bound by shared-memory fill bandwidth when every tap re-loads its own shifted x tile.
//
// Per 16 x 8 pixel block the CTA loads ONE (18 x 10) halo patch of x and ONE dy tile.  The GEMM K axis is
// the pixel axis (MN-major operands), a K step of 16 pixels = two 8-pixel tile rows = two 8-row groups
// SBO = 10 halo rows apart.  The 128 accumulator rows are `128 / Cin` chunks LBO = ONE pixel row apart,
// i.e. chunk c is the same patch shifted by c pixels: one tcgen05.mma computes the taps (dh, dw = c) for
// c = 0.. at once (chunks beyond dw = 2 are don't-care rows).  Accumulators: one per (dh, chunk group).
// =====================================================================================================
struct WhParams {
  int N, H, W, cin, cout, JT, KSB;
  int cs, nslab;               // channel slab of x handled by one task (16 / 32 / 64), cin / cs slabs
  int nacc_per_dh;             // 1 (slab <= 32: dw 0..2 in one MMA) or 2 (slab = 64: [dw0,dw1] and [dw2,-])
  int nsplit, stages;
  float* dw;
  // dh-merged variant: the three tap ROWS sit side by side in N (N = 3 * Cout): the dy tile is loaded with a one-row
  // halo above and below (18 x 8) and N chunk c of the B descriptor starts c tile rows further down (LBO = one tile row),
  // i.e. it pairs the x pixel q with dy[q - (dh, 0)], dh = 1 - c; the x patch then needs its dw halo only (16 x 10).
  // One MMA instead of three per K step: 32 + N/4 cycles each (tools/mma_probe.cu) -> 56 instead of 3 * 40 at Cout = 32.
  int dhm;
};
struct WhMaps { CUtensorMap a, b; };

__global__ void __launch_bounds__(kThreadsW, 2) tc_wgrad_halo_kernel(const __grid_constant__ WhMaps maps,
                                                                      const __grid_constant__ WhParams prm) {
  B2U_PDL_LAUNCH_DEPENDENTS();      // B2U_PDL_WAIT() follows the CTA-local setup (barriers, TMEM)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int cin = prm.cs, JT = prm.JT, KSB = prm.KSB, stages = prm.stages;   // `cin` = slab width from here on
  const uint32_t rowa = cin * 2, rowb = KSB * 2;
  const int dhm = prm.dhm;
  const uint32_t a_bytes = (dhm ? 160u : 180u) * rowa, a_stride = (a_bytes + 1023) & ~1023u;
  const int nb = JT / KSB;                                      // (dhm: KSB == JT, one chunk per tap row)
  const uint32_t b_tile = (dhm ? 144u : 128u) * rowb, b_bytes = b_tile * nb;
  const uint32_t stage_stride = a_stride + b_bytes;
  uint8_t* tail = smem + stages * stage_stride;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (role dispatch)
  const int lane = threadIdx.x & 31;
  const int tiles_w = (prm.W + 7) / 8, tiles_h = (prm.H + 15) / 16;
  const int nblocks = prm.N * tiles_h * tiles_w;
  const int nj = prm.cout / JT;
  const int njs = nj * prm.nslab;                             // task = (split, slab, jt), jt fastest
  const int ntasks = prm.nsplit * njs;
  const int nacc = dhm ? prm.nacc_per_dh : 3 * prm.nacc_per_dh;
  const int NW = dhm ? 3 * JT : JT;                             // accumulator width = MMA N
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(nacc * NW)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(tfull_bar, 1);
    tc::mbar_init(tempty_bar, 4);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, tmem_cols);
  B2U_PDL_WAIT();                    // everything below may read what the preceding kernel wrote
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int jt = task % nj, slab = (task / nj) % prm.nslab, sp = task / njs;
        for (int blk = sp; blk < nblocks; blk += prm.nsplit) {
          const int tw = blk % tiles_w, th = (blk / tiles_w) % tiles_h, n = blk / (tiles_w * tiles_h);
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_stride;
          tc::mbar_expect_tx(&full_bar[stage], a_bytes + b_bytes);
          tc::tma_load_4d(sa, &maps.a, &full_bar[stage], slab * cin, tw * 8 - 1, th * 16 - (dhm ? 0 : 1), n);
          for (int s = 0; s < nb; ++s)
            tc::tma_load_4d(sa + a_stride + s * b_tile, &maps.b, &full_bar[stage], jt * JT + s * KSB, tw * 8,
                            th * 16 - (dhm ? 1 : 0), n);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, uniform control flow; one elected lane issues
      const uint32_t idesc = tc::idesc_f16(128, NW, 1, 1);
      const uint64_t la = cin == 64 ? tc::SWZ_128B : (cin == 32 ? tc::SWZ_64B : tc::SWZ_32B);
      const uint64_t lb = KSB == 64 ? tc::SWZ_128B : (KSB == 32 ? tc::SWZ_64B : tc::SWZ_32B);
      // A: chunks one pixel row apart (LBO = rowa), 8-row K groups 10 halo rows apart (SBO = 10 * rowa)
      const uint32_t a_hi = (uint32_t)(tc::smem_desc(0, rowa, 10 * rowa, la) >> 32);
      const uint32_t b_hi = (uint32_t)(tc::smem_desc(0, b_tile, 8 * rowb, lb) >> 32);
      // B chunk stride: the next KSB-wide tile (JT > KSB), or -- dh-merged -- the next dy tile row (8 pixels)
      const uint32_t a_lbo = ((rowa >> 4) & 0x3FFF) << 16, b_lbo = ((((dhm ? 8u * rowb : b_tile)) >> 4) & 0x3FFF) << 16;
      const uint32_t smem_lo = (tc::smem_u32(smem) & 0x3FFFF) >> 4;
      const uint32_t stage16 = stage_stride >> 4, a_stride16 = a_stride >> 4;
      const uint32_t rowa16 = rowa >> 4, kb16 = (16 * rowb) >> 4;
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int sp = task / njs;
        tc::mbar_wait(tempty_bar, tphase ^ 1);
        tc::fence_after_sync();
        uint32_t first = 1;
        for (int blk = sp; blk < nblocks; blk += prm.nsplit) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::fence_after_sync();
          const uint32_t a0 = smem_lo + (uint32_t)stage * stage16;
          const uint32_t b0 = (a0 + a_stride16) | b_lbo;
          if (dhm) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b0 + kk * kb16);
              const uint32_t ar = (a0 + (uint32_t)(2 * kk * 10) * rowa16) | a_lbo;
              const uint32_t accf = (kk != 0) ? 1u : (first ? 0u : 1u);
              tc::mma_f16_ss_elect(tmem_base, ((uint64_t)a_hi << 32) | (uint64_t)ar, bd, idesc, accf);
              if (prm.nacc_per_dh == 2)
                tc::mma_f16_ss_elect(tmem_base + NW, ((uint64_t)a_hi << 32) | (uint64_t)(ar + 2 * rowa16), bd, idesc, accf);
            }
          } else {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b0 + kk * kb16);
#pragma unroll
            for (int dh = 0; dh < 3; ++dh) {
              // halo row of the first pixel of this K step for tap row dh: (2*kk + dh) * 10
              const uint32_t ar = (a0 + (uint32_t)((2 * kk + dh) * 10) * rowa16) | a_lbo;
              const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)ar;
              tc::mma_f16_ss_elect(tmem_base + (dh * prm.nacc_per_dh) * JT, ad, bd, idesc, (kk != 0) ? 1u : (first ? 0u : 1u));
              if (prm.nacc_per_dh == 2) {
                const uint64_t ad2 = ((uint64_t)a_hi << 32) | (uint64_t)(ar + 2 * rowa16);
                tc::mma_f16_ss_elect(tmem_base + (dh * 2 + 1) * JT, ad2, bd, idesc, (kk != 0) ? 1u : (first ? 0u : 1u));
              }
            }
          }
          }
          first = 0;
          tc::mma_commit_elect(&empty_bar[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        tc::mma_commit_elect(tfull_bar);
        tphase ^= 1;
      }
    }
  } else {
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int chunk = row / cin, ci = row % cin;              // chunk = dw offset inside an accumulator
    uint32_t tphase = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
      const int jt = task % nj, slab = (task / nj) % prm.nslab;
      tc::mbar_wait(tfull_bar, tphase);
      tc::fence_after_sync();
      for (int a = 0; a < nacc; ++a) {
        const int dwp = chunk + (prm.nacc_per_dh == 2 ? 2 * (a % 2) : 0);
        const bool live = dwp < 3;
        for (int cw = 0; cw < NW; cw += 16) {
          // dh-merged: N chunk cw / JT holds tap row 2 - chunk (see WhParams::dhm)
          const int dh = dhm ? 2 - cw / JT : a / prm.nacc_per_dh;
          const int c0 = dhm ? cw % JT : cw;
          float* out = prm.dw + ((long long)((dh * 3 + (live ? dwp : 0)) * prm.cin + slab * cin + ci)) * prm.cout + jt * JT;
          float v[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + a * NW + cw, v);
          if (live) {
// 16-byte vector reductions (red.global.add.v4.f32): 4 L2 atomic operations instead of 16
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              atomicAdd(reinterpret_cast<float4*>(out + c0 + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tempty_bar);
      tphase ^= 1;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;

int get_enc() {
  if (g_enc != nullptr) return B2U_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B2U_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    b2u_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B2U_ERR_CUDA;
  }
  g_enc = (EncodeTiledFn)fn;
  return B2U_OK;
}

int act_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sW, long long sH, long long sN,
            int KS) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)KS, (cuuint32_t)kTileW, (cuuint32_t)kTileH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = KS == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KS == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b2u_set_error("cuTensorMapEncodeTiled(wgrad C=%d W=%d H=%d N=%d KS=%d) failed: %d", C, W, H, N, KS, (int)r);
    return B2U_ERR_CUDA;
  }
  return B2U_OK;
}

int ks_for(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : (c % 16 == 0 ? 16 : 0)); }

bool g_attr = false;

int launch_wg(const WgMaps& maps, WgParams& p, void* stream) {
  const size_t a_bytes = 128 * 128 * 2, b_bytes = (size_t)p.JT * 256;
  p.stages = (int)((200 * 1024) / (a_bytes + b_bytes));
  if (p.stages > 6) p.stages = 6;
  B2U_REQUIRE(p.stages >= 2, "tc_wgrad: tile too large for a 2-stage pipeline");
  size_t smem = 1024 + p.stages * (a_bytes + b_bytes) + 256;
  const int nblocks = p.N * b2u_cdiv(p.H, kTileH) * b2u_cdiv(p.W, kTileW);
  const int base = p.ngroups * (p.J / p.JT);
  int ns = (2 * B2U_NUM_SMS + base - 1) / base;              // ~2 tasks per SM
  if (ns > nblocks) ns = nblocks;
  if (ns < 1) ns = 1;
  p.nsplit = ns;
  if (!g_attr) {
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_attr = true;
  }
  int ntasks = ns * base;
  int grid = ntasks < B2U_NUM_SMS ? ntasks : B2U_NUM_SMS;
  B2U_LAUNCH(tc_wgrad_kernel, grid, kThreadsW, smem, stream, maps, p);
  return B2U_OK;
}

int pick_jt_w(int J) {            // N tile: <= 128 columns, multiple of the slab width
  if (J <= 128) return J;
  return J % 128 == 0 ? 128 : (J % 64 == 0 ? 64 : (J % 32 == 0 ? 32 : 16));
}

}  // namespace

int b2u_channel_sum_f16(const void* dy, int lddy, int c, long long npix, float* db, void* stream);

int b2u_tc_wgrad_ok(int cin, int cout, int ldx, int lddy) {
  if (ks_for(cin) == 0 || ks_for(cout) == 0 || ldx % 8 || lddy % 8 || cout % 16) return 0;
  // row groups pack (tap, channel-slab) chunks of KSA channels, 128 rows per group; any Cin that is a multiple of
  // its slab width works (96 -> 3 slabs of 32, 192 -> 3 slabs of 64: U-Net++ concat widths)
  const int ksa = cin >= 128 ? 64 : ks_for(cin);
  if (cin % ksa) return 0;
  const int chunks = 9 * (cin / ksa), per_group = 128 / ksa;
  return (chunks + per_group - 1) / per_group <= kMaxGroups;
}
int b2u_tc_convt_wgrad_ok(int cin, int cout, int ldx, int lddy) {
  if (ks_for(cin) == 0 || ks_for(cout) == 0 || ldx % 8 || lddy % 8 || cin % 16) return 0;
  if (cout >= 128 && cout % 128) return 0;
  if (cout < 128 && (128 % cout || 128 / cout > 4)) return 0;
  const int groups = cout >= 128 ? 4 * (cout / 128) : (4 * cout + 127) / 128;
  return groups <= kMaxGroups;
}

// dw[t][ci][co] += sum_p x[p + off_t][ci] * dy[p][co];  db[co] += sum_p dy[p][co]
int g_b2u_wgrad_dhm = 1;    // 1: dh-merged tiles for Cout = 32 / 64 (see WhParams::dhm)
int g_b2u_wgrad_halo = 2;   // 0: per-tap tiles only, 1: halo for Cin <= 64, 2: + 32-channel slabs for wide layers, 3: 64-channel slabs
bool g_attr_h = false;

// Split-K factor: tasks = ns * base are dealt round-robin to `ctas` persistent CTAs, so the last round is
// only full when ns * base is close to a multiple of ctas; each task should still own >= ~6 pixel blocks so
// that its pipeline fill and its accumulator flush (fp32 atomics) stay amortised.
static int pick_nsplit(int base, int ctas, int nblocks) {
  int best = 1;
  double best_score = -1.0;
  for (int ns = 1; ns <= nblocks && ns <= 64; ++ns) {
    const long long tasks = (long long)ns * base;
    const long long rounds = (tasks + ctas - 1) / ctas;
    const double eff = (double)tasks / (double)(rounds * ctas);
    const int blocks_per_task = nblocks / ns;
    if (ns > 1 && blocks_per_task < 6) break;
    // per-task fixed cost ~ 2 pixel blocks (fill + flush)
    const double score = eff * (double)blocks_per_task / (double)(blocks_per_task + 2);
    if (score > best_score + 1e-9) { best_score = score; best = ns; }
  }
  return best;
}

static int wgrad_halo(const void* x, int ldx, int cin, int cs, const void* dy, int lddy, int cout, float* dw, int n, int h,
                      int wd, void* stream, int dhm = 0) {
  WhParams p{};
  p.N = n; p.H = h; p.W = wd; p.cin = cin; p.cout = cout; p.dw = dw;
  p.cs = cs; p.nslab = cin / cs;
  p.nacc_per_dh = cs == 64 ? 2 : 1;
  p.dhm = dhm;
  const int nacc = dhm ? p.nacc_per_dh : 3 * p.nacc_per_dh;
  int jt = cout <= 128 ? cout : 128;
  if (dhm) {
    B2U_REQUIRE((cout == 32 || cout == 64) && nacc * 3 * cout <= 256, "tc_wgrad_halo: dh-merged tiles need Cout 32 / 64 (cout=%d cs=%d)",
                cout, cs);
  } else {
    while (nacc * jt > 512 || cout % jt) jt -= 16;
  }
  B2U_REQUIRE(jt >= 16, "tc_wgrad_halo: no N tile for cout=%d", cout);
  p.JT = jt;
  p.KSB = ks_for(jt);
  const size_t a_stride = ((size_t)(dhm ? 160 : 180) * cs * 2 + 1023) & ~(size_t)1023;
  const size_t b_bytes = dhm ? (size_t)144 * jt * 2 : (size_t)jt * 256;
  const size_t tailb = 256;
  // two CTAs per SM (two MMA issuers) when both fit TMEM (512 columns per SM) -- else one CTA with a deep ring
  int cols = 32;
  while (cols < nacc * (dhm ? 3 * jt : jt)) cols <<= 1;
  const int per_sm = cols <= 256 ? 2 : 1;
  const size_t cap = per_sm == 2 ? 108 * 1024 : 216 * 1024;
  int st = (int)((cap - 1024 - tailb) / (a_stride + b_bytes));
  if (st > 8) st = 8;
  B2U_REQUIRE(st >= 2, "tc_wgrad_halo: tiles do not fit (cin=%d cout=%d)", cin, cout);
  p.stages = st;
  const size_t smem = 1024 + (size_t)st * (a_stride + b_bytes) + tailb;
  const int nblocks = n * b2u_cdiv(h, 16) * b2u_cdiv(wd, 8);
  const int nj = cout / jt;
  const int base = nj * p.nslab;
  int ns;
  if (p.nslab == 1) {
    ns = (per_sm * B2U_NUM_SMS + nj - 1) / nj;
    if (ns > nblocks) ns = nblocks;
    if (ns < 1) ns = 1;
  } else {
    ns = pick_nsplit(base, per_sm * B2U_NUM_SMS, nblocks);
  }
  p.nsplit = ns;
  WhMaps maps;
  {
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)wd * ldx * 2, (cuuint64_t)h * wd * ldx * 2};
    cuuint32_t box[4] = {(cuuint32_t)cs, 10, (cuuint32_t)(dhm ? 16 : 18), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMapSwizzle sw = cs == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cs == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = g_enc(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b2u_set_error("tc_wgrad_halo: x tensor map failed (%d)", (int)r); return B2U_ERR_CUDA; }
    cuuint64_t bd[4] = {(cuuint64_t)cout, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t bs[3] = {(cuuint64_t)lddy * 2, (cuuint64_t)wd * lddy * 2, (cuuint64_t)h * wd * lddy * 2};
    cuuint32_t bb[4] = {(cuuint32_t)p.KSB, 8, (cuuint32_t)(dhm ? 18 : 16), 1};
    CUtensorMapSwizzle swb = p.KSB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.KSB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    r = g_enc(&maps.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(dy), bd, bs, bb, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swb, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b2u_set_error("tc_wgrad_halo: dy tensor map failed (%d)", (int)r); return B2U_ERR_CUDA; }
  }
  if (!g_attr_h) {
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_attr_h = true;
  }
  const int ntasks = ns * base;
  const int grid = ntasks < per_sm * B2U_NUM_SMS ? ntasks : per_sm * B2U_NUM_SMS;
  B2U_LAUNCH(tc_wgrad_halo_kernel, grid, kThreadsW, smem, stream, maps, p);
  return B2U_OK;
}

int b2u_tc_conv3x3_wgrad(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw, float* db,
                         int n, int h, int wd, void* ws, size_t ws_bytes, void* stream) {
  (void)ws; (void)ws_bytes;
  int rc = get_enc();
  if (rc != B2U_OK) return rc;
  // halo variant: thin layers take the whole Cin as one slab; wide layers (wgrad_halo >= 2) are cut into 32-channel
  // slabs (64-channel with wgrad_halo == 3), one slab per task -- 4x less L2 -> shared-memory traffic than the
  // per-tap tiles below, which is what bounds them (42 B/clk/SM of L2 bandwidth against 128 B/clk of operands)
  int cs = 0;
  if (g_b2u_wgrad_halo && (cin == 16 || cin == 32 || cin == 64)) cs = (g_b2u_wgrad_halo == 4 && cin == 64) ? 32 : cin;
  else if (g_b2u_wgrad_halo >= 2 && cin > 64 && cin % 32 == 0) cs = (g_b2u_wgrad_halo == 3 && cin % 64 == 0) ? 64 : 32;
  // thin outputs (Cout = 32 / 64): the three tap rows merged in N (WhParams::dhm); its accumulators (3 * Cout columns per
  // 128 rows) must leave room for two CTAs per SM, so 64-channel slabs only with Cout = 32
  int dhm = 0;
  if (g_b2u_wgrad_dhm && g_b2u_wgrad_halo && (cout == 32 || cout == 64) && cin % 32 == 0) {
    dhm = 1;
    cs = (cin == 64 && cout == 32) ? 64 : 32;
  }
  if (cs != 0 && cout % 16 == 0) {
    rc = wgrad_halo(x, ldx, cin, cs, dy, lddy, cout, dw, n, h, wd, stream, dhm);
    if (rc != B2U_OK) return rc;
    if (db != nullptr) return b2u_channel_sum_f16(dy, lddy, cout, (long long)n * h * wd, db, stream);
    return B2U_OK;
  }
  WgParams p{};
  p.N = n; p.H = h; p.W = wd;
  p.KSA = cin >= 128 ? 64 : ks_for(cin);
  if (cin < 128 && p.KSA != cin) p.KSA = ks_for(cin);        // e.g. cin = 96 -> 32-wide chunks
  p.nchunks = 128 / p.KSA;
  p.KSB = ks_for(cout); p.J = cout; p.JT = pick_jt_w(cout);
  if (p.JT % p.KSB) p.KSB = ks_for(p.JT);
  p.out_row_stride = cout; p.dw = dw;
  // row groups: each chunk is (tap, channel slab); rows of a chunk are consecutive input channels
  const int slabs = cin / p.KSA;                              // chunks per tap
  int g = 0, c = 0;
  for (int t = 0; t < 9; ++t) {
    for (int s = 0; s < slabs; ++s) {
      B2U_REQUIRE(g < kMaxGroups, "tc_wgrad: too many row groups (cin=%d)", cin);
      p.ch_map[g][c] = 0; p.ch_dh[g][c] = (int8_t)(t / 3 - 1); p.ch_dw[g][c] = (int8_t)(t % 3 - 1);
      p.ch_c0[g][c] = (int16_t)(s * p.KSA);
      p.out_base[g][c] = (t * cin + s * p.KSA) * cout;
      if (++c == p.nchunks) { c = 0; ++g; }
    }
  }
  if (c != 0) {
    for (; c < p.nchunks; ++c) p.ch_map[g][c] = -1;
    ++g;
  }
  p.ngroups = g;
  WgMaps maps;
  rc = act_map(&maps.a[0], x, cin, wd, h, n, ldx, (long long)wd * ldx, (long long)h * wd * ldx, p.KSA);
  if (rc != B2U_OK) return rc;
  for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
  rc = act_map(&maps.b, dy, cout, wd, h, n, lddy, (long long)wd * lddy, (long long)h * wd * lddy, p.KSB);
  if (rc != B2U_OK) return rc;
  rc = launch_wg(maps, p, stream);
  if (rc != B2U_OK) return rc;
  if (db != nullptr) return b2u_channel_sum_f16(dy, lddy, cout, (long long)n * h * wd, db, stream);
  return B2U_OK;
}

// dw[ab][co][ci] += sum_p dy[n,2i+a,2j+b,co] * x[n,i,j,ci];  db[co] += sum over all output pixels of dy
int b2u_tc_convt_wgrad(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw, float* db, int n,
                       int h, int wd, void* ws, size_t ws_bytes, void* stream) {
  (void)ws; (void)ws_bytes;
  int rc = get_enc();
  if (rc != B2U_OK) return rc;
  WgParams p{};
  p.N = n; p.H = h; p.W = wd;
  p.KSA = cout >= 128 ? 64 : ks_for(cout);
  p.nchunks = 128 / p.KSA;
  p.KSB = ks_for(cin); p.J = cin; p.JT = pick_jt_w(cin);
  if (p.JT % p.KSB) p.KSB = ks_for(p.JT);
  p.out_row_stride = cin; p.dw = dw;
  const int slabs = cout / p.KSA;
  int g = 0, c = 0;
  for (int ab = 0; ab < 4; ++ab) {
    for (int s = 0; s < slabs; ++s) {
      B2U_REQUIRE(g < kMaxGroups, "tc_convt_wgrad: too many row groups (cout=%d)", cout);
      p.ch_map[g][c] = (int8_t)ab; p.ch_dh[g][c] = 0; p.ch_dw[g][c] = 0;
      p.ch_c0[g][c] = (int16_t)(s * p.KSA);
      p.out_base[g][c] = (ab * cout + s * p.KSA) * cin;
      if (++c == p.nchunks) { c = 0; ++g; }
    }
  }
  if (c != 0) {
    for (; c < p.nchunks; ++c) p.ch_map[g][c] = -1;
    ++g;
  }
  p.ngroups = g;
  WgMaps maps;
  for (int ab = 0; ab < 4; ++ab) {
    const __half* base = (const __half*)dy + ((long long)(ab >> 1) * (2 * wd) + (ab & 1)) * lddy;
    rc = act_map(&maps.a[ab], base, cout, wd, h, n, 2LL * lddy, 4LL * wd * lddy, 4LL * h * wd * lddy, p.KSA);
    if (rc != B2U_OK) return rc;
  }
  rc = act_map(&maps.b, x, cin, wd, h, n, ldx, (long long)wd * ldx, (long long)h * wd * ldx, p.KSB);
  if (rc != B2U_OK) return rc;
  rc = launch_wg(maps, p, stream);
  if (rc != B2U_OK) return rc;
  if (db != nullptr) return b2u_channel_sum_f16(dy, lddy, cout, 4LL * n * h * wd, db, stream);
  return B2U_OK;
}
