// tcgen05 3x3 convolution with HALO-TILE reuse (forward and data gradient).
//
// conv_tc.cu loads one activation box per filter tap: the same pixels cross L2 -> shared memory nine
// times, which makes the thin, full-resolution layers (32/64 channels at 512^2 / 256^2: the HBM-bound
// half of the U-Net) L2- and TMA-latency-bound.  Here a CTA loads the (16+2) x (8+2) pixel halo patch of
// its 16 x 8 output tile ONCE per 64-channel slab and issues the nine taps as nine tcgen05.mma groups
// whose A descriptors start at row (dh*10 + dw) of that patch: a tile row (8 pixels) is one 8-row
// core-matrix group, consecutive tile rows are 10 halo rows apart (SBO = 10 rows).  The 128B/64B/32B
// swizzle applied by TMA is a function of the shared-memory address bits, so a row-shifted start stays
// consistent (`amode` 1; `amode` 3 is the conservative variant: three boxes, one per dw, whose tap starts
// are whole 8-row groups).  Weights of thin layers (9*K*J*2 bytes <= ~110 KB) are loaded once and stay
// RESIDENT in shared memory for the whole persistent CTA; wide layers stream per-tap weight tiles through
// a second ring.
//
// Warp roles: TMA producer, single-thread MMA issuer, 8 epilogue warps (the epilogue is instruction-latency
// bound with 4: one warp per scheduler cannot hide its own dependent-issue latency),
// double-buffered TMEM accumulator, bias / ReLU / ELU / mask / accumulate / BN statistics).
#include <cuda.h>
#include <type_traits>
#include "common.cuh"
#include "internal.h"
#include "launch.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kTH = 16, kTW = 8;            // output tile: 16 rows x 8 columns = 128 pixels
constexpr int kMaxEpiWarps = 8;             // 8: two warps per TMEM lane group (column halves); 4 when two CTAs share an SM
constexpr int kThreads3 = 64 + 32 * kMaxEpiWarps;
constexpr int kMaxSA = 8, kMaxSB = 8;

struct C3Params {
  int N, H, W, K, J, KS, JT;
  int amode;              // 1: one halo box (18 x 10), 3: three boxes (18 x 8), one per dw
  int bo_mode;            // amode 1: 0 = descriptor base_offset 0, 1 = (start >> 7) & 7
  int bres;               // weights resident in smem
  int SA, SB;             // ring depths (SB counts GROUPS of `bgroup` weight tiles)
  int mcast;              // streamed weights, launched as clusters of two CTAs: both CTAs work on the same N tile and the
                          // same K order for two different pixel tiles, and every weight tile is loaded ONCE for the pair
                          // (TMA multicast, the two CTAs alternate as the loader): half the L2 -> SM weight traffic that
                          // bounds the wide layers (profiles/NOTES_r2.md)
  int bgroup;             // streamed weights: tiles per ring slot (1, or 3 = the taps of one filter row).  A
                          // tcgen05.commit makes the issuing thread's next MMA start >= ~465 cycles after the previous
                          // batch's first (tools/mma_probe.cu): four N = 128 MMAs (256 cycles) per commit ran at 55 % of the
                          // pipe rate, twelve (768 cycles) do not notice it.  N = 256 tiles (512 cycles per four) keep 1.
  int epi_warps;          // 4 or 8 (block = 64 + 32 * epi_warps threads)
  uint32_t a_sub;         // bytes of one A sub-tile (1024-aligned), a_stage = a_sub * (amode == 3 ? 3 : 1)
  __half* y; int ldy;
  const float* bias; int act;
  const __half* mask; int ldmask; int mask_act;
  int accumulate;
  double* stats;          // BatchNorm statistics of the stored values: sums at [c], squares at [J + c]
  float* colsum;          // per-channel sums of the stored values, added with fp32 atomics (bias gradient of the
                          // layer whose output gradient this kernel writes)
  long long* dbg;         // optional timeline buffer (CTA 0): [iter][8] clock64 stamps
  // (appended last: existing constant-bank offsets stay as they were)
  uint8_t* bits_out;      // kF_BITS_OUT: packed 1-bit ReLU mask of the values stored (bit pix*J + column)
  const uint8_t* bits_in; // kF_BITS_IN : packed 1-bit ReLU mask applied to the values written (data gradient)
  const float* post_scale;  // kF_POST: y = post_scale[c] * act(conv + bias) + post_shift[c] -- the inference-mode
  const float* post_shift;  // BatchNormalization that follows the activation (T2:749-750), folded into the epilogue
};

struct C3Maps {
  CUtensorMap a;          // activations, box (KS, 10 | 8, 18, 1)
  CUtensorMap b;          // packed weights [9][J][K], box (KS, JT, 1)
};

__device__ __forceinline__ float transpose_reduce16_(float v[16], int lane) {
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

// kFlags specialises the epilogue at compile time (the kernel sits at the 168-register limit of its launch
// bounds and its thin-layer speed moves by 20 % with register allocation, so features a launch does not use must not
// cost it anything): bit 0 = activation-derivative mask and/or accumulate (data gradients), bit 1 = BatchNorm
// statistics and/or column sums.
constexpr int kF_MASKACC = 1, kF_SUMS = 2;
constexpr int kF_BITS_OUT = 4, kF_BITS_IN = 8;
constexpr int kF_POST = 16;      // per-channel affine after the activation (never together with statistics: the two
                                 // share the s_stats area of shared memory)
template <int kFlags>
__global__ void __launch_bounds__(kThreads3, 1) tc_conv3_kernel(const __grid_constant__ C3Maps maps,
                                                                 const __grid_constant__ C3Params prm) {
  constexpr bool kMaskAcc = (kFlags & kF_MASKACC) != 0, kSums = (kFlags & kF_SUMS) != 0;
  B2U_PDL_LAUNCH_DEPENDENTS();      // B2U_PDL_WAIT() follows the CTA-local setup (barriers, TMEM)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KS = prm.KS, JT = prm.JT, SA = prm.SA, SB = prm.SB;
  const int kslabs = prm.K / KS;
  const uint32_t rowb = KS * 2;                                   // bytes per pixel row
  const uint32_t a_stage = prm.a_sub * (prm.amode == 3 ? 3 : 1);
  const uint32_t b_tile = ((uint32_t)JT * rowb + 1023) & ~1023u;  // one (tap, slab) weight tile
  // layout: [A ring][B ring or resident weights][barriers][tmem ptr][bias][stats]
  uint8_t* a_ring = smem;
  uint8_t* b_area = a_ring + (size_t)SA * a_stage;
  const uint32_t b_bytes_total = prm.bres ? 9u * kslabs * b_tile : (uint32_t)(SB * prm.bgroup) * b_tile;
  uint8_t* tail = b_area + b_bytes_total;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* a_empty = a_full + kMaxSA;
  uint64_t* b_full = a_empty + kMaxSA;
  uint64_t* b_empty = b_full + kMaxSB;
  uint64_t* w_full = b_empty + kMaxSB;
  uint64_t* tfull = w_full + 1;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr + 4) + 15) & ~(uintptr_t)15);
  float* s_stats = s_bias + prm.J;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (role dispatch)
  const int lane = threadIdx.x & 31;
  const int tiles_w = (prm.W + kTW - 1) / kTW, tiles_h = (prm.H + kTH - 1) / kTH;
  const int nj = prm.J / JT;
  const int ntiles = prm.N * tiles_h * tiles_w * nj;          // host guarantees < 2^31
  // tile walk: CTA b takes tiles b, b + grid, ... (N tile fastest); clusters of two (mcast) walk "super tiles" = one N tile
  // x a PAIR of pixel tiles, CTA rank r taking pixel tile 2 * pair + r (the last pair of an odd count has a dummy tile: its
  // loads are zero-filled by TMA and nothing is stored)
  const uint32_t crank = prm.mcast ? tc::cluster_ctarank() : 0u;
  const unsigned CS = prm.mcast ? (unsigned)prm.mcast : 1u;          // cluster size (2, 4 or 8)
  const unsigned npt = (unsigned)(prm.N * tiles_h * tiles_w);
  const unsigned t_first = blockIdx.x / CS;
  const unsigned t_step = gridDim.x / CS;
  const unsigned t_end = prm.mcast ? ((npt + CS - 1) / CS) * (unsigned)nj : (unsigned)ntiles;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * JT)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], prm.mcast ? prm.mcast : 1); }
    tc::mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], prm.epi_warps); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, tmem_cols);
  B2U_PDL_WAIT();                    // everything below may read what the preceding kernel wrote
  for (int i = threadIdx.x; i < prm.J; i += blockDim.x) s_bias[i] = prm.bias ? prm.bias[i] : 0.f;
  if constexpr ((kFlags & kF_POST) != 0) {
    for (int i = threadIdx.x; i < prm.J; i += blockDim.x) {
      s_stats[i] = prm.post_scale[i];
      s_stats[prm.J + i] = prm.post_shift[i];
    }
  } else {
    for (int i = threadIdx.x; i < 2 * prm.J; i += blockDim.x) s_stats[i] = 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (prm.mcast) tc::cluster_sync();          // the peer's barriers are initialised before anything is multicast to them
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer =========================================
    if (lane == 0) {
      tc::prefetch_tmap(&maps.a);
      tc::prefetch_tmap(&maps.b);
      if (prm.bres) {
        tc::mbar_expect_tx(w_full, 9u * kslabs * (uint32_t)JT * rowb);
        for (int t = 0; t < 9; ++t)
          for (int ks = 0; ks < kslabs; ++ks)
            tc::tma_load_3d(b_area + (size_t)(t * kslabs + ks) * b_tile, &maps.b, w_full, ks * KS, 0, t);
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const uint32_t a_tx = (prm.amode == 3 ? 3u * 18 * 8 : 18u * 10) * rowb;
      unsigned slot_no = 0;                          // weight ring slots issued so far (mcast: parity = loading CTA)
      for (unsigned tile = t_first; tile < t_end; tile += t_step) {
        const int jt = (int)(tile % (unsigned)nj);
        const unsigned pt = CS * (tile / (unsigned)nj) + crank;
        const int tw = (int)(pt % (unsigned)tiles_w);
        const unsigned r = pt / (unsigned)tiles_w;
        const int th = (int)(r % (unsigned)tiles_h), n = (int)(r / (unsigned)tiles_h);
        for (int ks = 0; ks < kslabs; ++ks) {
          tc::mbar_wait(&a_empty[sa], pa ^ 1);
          if (prm.dbg && blockIdx.x == 0 && tile / t_step < 64) prm.dbg[(tile / t_step) * 8 + 0] = clock64();
          uint8_t* dst = a_ring + (size_t)sa * a_stage;
          tc::mbar_expect_tx(&a_full[sa], a_tx);
          if (prm.amode == 3) {
            for (int dw = 0; dw < 3; ++dw)
              tc::tma_load_4d(dst + dw * prm.a_sub, &maps.a, &a_full[sa], ks * KS, tw * kTW - 1 + dw, th * kTH - 1, n);
          } else {
            tc::tma_load_4d(dst, &maps.a, &a_full[sa], ks * KS, tw * kTW - 1, th * kTH - 1, n);
          }
          if (prm.dbg && blockIdx.x == 0 && tile / t_step < 64) prm.dbg[(tile / t_step) * 8 + 1] = clock64();
          if (++sa == SA) { sa = 0; pa ^= 1; }
          if (!prm.bres) {
            const int G = prm.bgroup;
            for (int t = 0; t < 9; t += G) {
              tc::mbar_wait(&b_empty[sb], pb ^ 1);        // (mcast: released by BOTH CTAs' MMA warps)
              tc::mbar_expect_tx(&b_full[sb], (uint32_t)(G * JT) * rowb);
              if (!prm.mcast) {
                for (int gi = 0; gi < G; ++gi)
                  tc::tma_load_3d(b_area + (size_t)(sb * G + gi) * b_tile, &maps.b, &b_full[sb], ks * KS, jt * JT, t + gi);
              } else if (slot_no % CS == crank) {          // this CTA's turn: one load fills the slot of every CTA
                for (int gi = 0; gi < G; ++gi)
                  tc::tma_load_3d_mcast(b_area + (size_t)(sb * G + gi) * b_tile, &maps.b, &b_full[sb], ks * KS, jt * JT,
                                        t + gi, (uint16_t)((1u << CS) - 1u));
              }
              ++slot_no;
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    // One thread issues every tcgen05.mma of the CTA.  It is a single dependent instruction stream, so the
    // descriptors are NOT rebuilt per MMA: the high words are constants, the low words (start address >> 4)
    // advance by precomputed offsets, and the 9-tap x K-step nest is fully unrolled (issue_slab<KSTEPS>).
    // The whole warp runs this loop in uniform control flow; one elected lane issues (see tc_common.cuh).
    {
      const uint32_t idesc = tc::idesc_f16(128, JT, 0, 0);
      const uint64_t layout = KS == 64 ? tc::SWZ_128B : (KS == 32 ? tc::SWZ_64B : tc::SWZ_32B);
      const uint32_t a_rows = prm.amode == 3 ? 8 : 10;            // halo row pitch in pixels
      // descriptor words: lo = (addr >> 4) | LBO(=1) << 16 ; hi = SBO >> 4 | version 1 << 14 | layout << 29
      const uint32_t a_hi = (uint32_t)(tc::smem_desc(0, 16, a_rows * rowb, layout) >> 32);
      const uint32_t b_hi = (uint32_t)(tc::smem_desc(0, 16, 8 * rowb, layout) >> 32);
      uint32_t tap_off[9];                                        // A start offset of tap t, in 16-byte units
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dh = t / 3, dw = t % 3;
        tap_off[t] = prm.amode == 3 ? (dw * prm.a_sub + (uint32_t)(dh * 8) * rowb) >> 4
                                    : ((uint32_t)(dh * 10 + dw) * rowb) >> 4;
      }
      const uint32_t a_ring_lo = ((tc::smem_u32(a_ring) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t b_area_lo = ((tc::smem_u32(b_area) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t a_stage16 = a_stage >> 4, b_tile16 = b_tile >> 4;
      if (prm.bres) { tc::mbar_wait(w_full, 0); tc::fence_after_sync(); }
      int sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, acc_phase = 0;
      auto issue_slab = [&](auto ksteps_tag, uint32_t d_tmem, uint32_t a_lo, int ks) {
        constexpr int KSTEPS = decltype(ksteps_tag)::value;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          uint32_t b_lo;
          const bool g3 = prm.bgroup == 3;
          if (prm.bres) {
            b_lo = b_area_lo + (uint32_t)(t * kslabs + ks) * b_tile16;
          } else {
            if (!g3 || t % 3 == 0) {                     // first tile of a ring slot
              tc::mbar_wait(&b_full[sb], pb);
              tc::fence_after_sync();
            }
            b_lo = b_area_lo + (uint32_t)(g3 ? sb * 3 + t % 3 : sb) * b_tile16;
          }
          const uint32_t at = a_lo + tap_off[t];
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk) {
            const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(at + 2 * kk);
            const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + 2 * kk);
            tc::mma_f16_ss_elect(d_tmem, ad, bd, idesc, (t | kk) != 0 ? 1u : (ks != 0 ? 1u : 0u));
          }
          if (!prm.bres && (!g3 || t % 3 == 2)) {        // last tile of the slot: one commit releases all of it
            if (prm.mcast) tc::mma_commit_mcast_elect(&b_empty[sb], (uint16_t)((1u << prm.mcast) - 1u));   // ... in every CTA
            else tc::mma_commit_elect(&b_empty[sb]);
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
        }
      };
      for (unsigned tile = t_first; tile < t_end; tile += t_step) {
        tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc::fence_after_sync();
        const bool dbg_on = prm.dbg && blockIdx.x == 0 && tile / t_step < 64;
        if (dbg_on && lane == 0) prm.dbg[(tile / t_step) * 8 + 2] = clock64();
        const uint32_t d_tmem = tmem_base + acc * JT;
        for (int ks = 0; ks < kslabs; ++ks) {
          tc::mbar_wait(&a_full[sa], pa);
          tc::fence_after_sync();
          if (dbg_on && ks == 0 && lane == 0) prm.dbg[(tile / t_step) * 8 + 3] = clock64();
          const uint32_t a_lo = a_ring_lo + (uint32_t)sa * a_stage16;
          if (KS == 64) issue_slab(std::integral_constant<int, 4>{}, d_tmem, a_lo, ks);
          else if (KS == 32) issue_slab(std::integral_constant<int, 2>{}, d_tmem, a_lo, ks);
          else issue_slab(std::integral_constant<int, 1>{}, d_tmem, a_lo, ks);
          tc::mma_commit_elect(&a_empty[sa]);
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
        tc::mma_commit_elect(&tfull[acc]);
        if (dbg_on && lane == 0) prm.dbg[(tile / t_step) * 8 + 4] = clock64();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue ==============================================
    // 8 warps: warp -> TMEM lane group (warp & 3) and column half ((warp - 2) >> 2).  A thread owns one
    // output pixel (accumulator row) and walks its columns 16 at a time.  BatchNorm statistics of thin
    // layers (<= 32 columns per thread) accumulate in registers across ALL tiles of the persistent CTA
    // and are reduced once at the end; wider layers reduce per tile with a shuffle transpose.
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int half = ew >> 2;
    const int row = lg * 32 + lane;
    const bool split = prm.epi_warps == 8 && JT % 32 == 0;          // two warps share a lane group's columns (halves of whole 16-column chunks)
    const int ccols = split ? JT / 2 : JT;                        // columns owned by this warp within a tile
    const int cbeg = split ? half * ccols : 0;
    const bool has_cols = split || half == 0;
    const bool want_sums = kSums && (prm.stats != nullptr || prm.colsum != nullptr);
    // register accumulators: 32 columns of (sum, sum of squares), or -- column sums only -- 64 columns of sums
    // (rs2 then holds columns 32..63)
    const bool wide_sums = prm.stats == nullptr && ccols > 32;
    const bool reg_stats = want_sums && nj == 1 && (ccols <= 32 || (wide_sums && ccols <= 64));
    constexpr int kRS = kSums ? 32 : 1;                            // register accumulators exist in kSums variants only
    constexpr int kR16 = kSums ? 16 : 0;
    float rs1[kRS], rs2[kRS];
#pragma unroll
    for (int i = 0; i < kRS; ++i) { rs1[i] = 0.f; rs2[i] = 0.f; }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (unsigned tile = t_first; tile < t_end; tile += t_step) {
      const int jt = (int)(tile % (unsigned)nj);
      const unsigned pt = CS * (tile / (unsigned)nj) + crank;
      const int tw = (int)(pt % (unsigned)tiles_w);
      const unsigned r = pt / (unsigned)tiles_w;
      const int th = (int)(r % (unsigned)tiles_h), n = (int)(r / (unsigned)tiles_h);
      const int h = th * kTH + (row >> 3), w = tw * kTW + (row & 7);
      const bool valid = h < prm.H && w < prm.W && pt < npt;
      const long long pix = ((long long)n * prm.H + h) * prm.W + w;
      __half* yrow = prm.y + pix * prm.ldy + jt * JT;
      const __half* mrow = (kMaskAcc && prm.mask != nullptr) ? prm.mask + pix * prm.ldmask + jt * JT : nullptr;
      // thin layers: issue the tile's activation-derivative mask loads BEFORE waiting for the accumulator, so that
      // their DRAM round trip overlaps the wait instead of following it chunk by chunk
      const bool mask_early = kMaskAcc && mrow != nullptr && ccols <= 32 && has_cols && valid;
      uint4 mk[4];
      if (mask_early) {
        const uint4* mp = reinterpret_cast<const uint4*>(mrow + cbeg);
        mk[0] = __ldg(mp);
        mk[1] = __ldg(mp + 1);
        if (ccols > 16) { mk[2] = __ldg(mp + 2); mk[3] = __ldg(mp + 3); }
      }
      // 1-bit ReLU mask of this thread's columns (<= 128): a few 2-byte loads issued before the accumulator wait
      // (`if constexpr`, and the bits live in mk[0] -- unused here, the fp16 mask path is off in these variants -- so
      // that the other variants compile to the code they had before this feature existed)
      if constexpr ((kFlags & kF_BITS_IN) != 0) {
        mk[0] = make_uint4(0u, 0u, 0u, 0u);
        if (has_cols && valid) {
          const unsigned short* bp16 =
              reinterpret_cast<const unsigned short*>(prm.bits_in + ((pix * prm.J + jt * JT + cbeg) >> 3));
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q * 16 < ccols) {
              const uint32_t hw = (uint32_t)__ldg(bp16 + q) << (16 * (q & 1));
              if (q < 2) mk[0].x |= hw; else if (q < 4) mk[0].y |= hw; else if (q < 6) mk[0].z |= hw; else mk[0].w |= hw;
            }
          }
        }
      }
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::fence_after_sync();
      const bool dbg_e = prm.dbg && blockIdx.x == 0 && tile / t_step < 64 && threadIdx.x == 64;
      if (dbg_e) prm.dbg[(tile / t_step) * 8 + 5] = clock64();
      if (has_cols) {
#pragma unroll 2
        for (int cc = 0; cc < ccols; cc += 16) {
          const int c0 = cbeg + cc;
          float v[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * JT + c0, v);
          const float4* bp = reinterpret_cast<const float4*>(s_bias + jt * JT + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = bp[q];
            v[4 * q + 0] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
          }
          // activation switch hoisted out of the element loop: the 16 element chains stay branch-free
          if (prm.act == B2U_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (prm.act == B2U_ACT_ELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
          }
          if constexpr ((kFlags & kF_POST) != 0) {
            const float4* sp = reinterpret_cast<const float4*>(s_stats + jt * JT + c0);
            const float4* tp = reinterpret_cast<const float4*>(s_stats + prm.J + jt * JT + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 s4 = sp[q], t4 = tp[q];
              v[4 * q + 0] = fmaf(v[4 * q + 0], s4.x, t4.x); v[4 * q + 1] = fmaf(v[4 * q + 1], s4.y, t4.y);
              v[4 * q + 2] = fmaf(v[4 * q + 2], s4.z, t4.z); v[4 * q + 3] = fmaf(v[4 * q + 3], s4.w, t4.w);
            }
          }
          if constexpr ((kFlags & kF_BITS_IN) != 0) {
            if (valid) {
              const uint32_t w32 = cc < 32 ? mk[0].x : (cc < 64 ? mk[0].y : (cc < 96 ? mk[0].z : mk[0].w));
              const uint32_t b16 = (w32 >> (cc & 16)) & 0xffffu;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = ((b16 >> i) & 1u) ? v[i] : 0.f;
            }
          }
          if (valid) {
            if (mask_early) {
              float m[8];
              unpack8h(cc == 0 ? mk[0] : mk[2], m);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= act_bwd_from_y(m[i], prm.mask_act);
              unpack8h(cc == 0 ? mk[1] : mk[3], m);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 + i] *= act_bwd_from_y(m[i], prm.mask_act);
            } else if (mrow != nullptr) {
              float m[8];
              load8<__half>(mrow + c0, m);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= act_bwd_from_y(m[i], prm.mask_act);
              load8<__half>(mrow + c0 + 8, m);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 + i] *= act_bwd_from_y(m[i], prm.mask_act);
            }
            if (kMaskAcc && prm.accumulate) {
              float e[8];
              load8<__half>(yrow + c0, e);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += e[i];
              load8<__half>(yrow + c0 + 8, e);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 + i] += e[i];
            }
            store8<__half>(yrow + c0, v);
            store8<__half>(yrow + c0 + 8, v + 8);
            if constexpr ((kFlags & kF_BITS_OUT) != 0) {
              // bit = (the fp16 value just stored > 0): round-to-nearest-even sends v <= 2^-25 to zero
              uint32_t b16 = 0u;
#pragma unroll
              for (int i = 0; i < 16; ++i) b16 |= (v[i] > 2.98023223876953125e-08f ? 1u : 0u) << i;
              *reinterpret_cast<unsigned short*>(prm.bits_out + ((pix * prm.J + jt * JT + c0) >> 3)) = (unsigned short)b16;
            }
          }
          if (want_sums) {
            if (reg_stats) {
              if (valid) {
                if (!wide_sums) {
                  if (cc == 0) {
#pragma unroll
                    for (int i = 0; i < kR16; ++i) { rs1[i] += v[i]; rs2[i] = fmaf(v[i], v[i], rs2[i]); }
                  } else {
#pragma unroll
                    for (int i = 0; i < kR16; ++i) { rs1[kR16 + i] += v[i]; rs2[kR16 + i] = fmaf(v[i], v[i], rs2[kR16 + i]); }
                  }
                } else if (cc == 0) {
#pragma unroll
                  for (int i = 0; i < kR16; ++i) rs1[i] += v[i];
                } else if (cc == 16) {
#pragma unroll
                  for (int i = 0; i < kR16; ++i) rs1[kR16 + i] += v[i];
                } else if (cc == 32) {
#pragma unroll
                  for (int i = 0; i < kR16; ++i) rs2[i] += v[i];
                } else {
#pragma unroll
                  for (int i = 0; i < kR16; ++i) rs2[kR16 + i] += v[i];
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
                float s2 = transpose_reduce16_(sq, lane);
                if (lane < 16) atomicAdd(&s_stats[prm.J + jt * JT + c0 + lane], s2);
              }
              float s1 = transpose_reduce16_(q, lane);
              if (lane < 16) atomicAdd(&s_stats[jt * JT + c0 + lane], s1);
            }
          }
        }
      }
      if (dbg_e) prm.dbg[(tile / t_step) * 8 + 6] = clock64();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (reg_stats && has_cols) {
      // register statistics hold columns [cbeg, cbeg + ccols) of the (single) N tile: J == JT here
      for (int cc = 0; cc < ccols; cc += 16) {
        float q[16], sq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int i0 = kSums ? i : 0, i1 = kSums ? 16 + i : 0;       // (never executed without kSums)
          if (!wide_sums) { q[i] = cc == 0 ? rs1[i0] : rs1[i1]; sq[i] = cc == 0 ? rs2[i0] : rs2[i1]; }
          else { q[i] = cc == 0 ? rs1[i0] : (cc == 16 ? rs1[i1] : (cc == 32 ? rs2[i0] : rs2[i1])); sq[i] = 0.f; }
        }
        float s1 = transpose_reduce16_(q, lane);
        if (lane < 16) atomicAdd(&s_stats[cbeg + cc + lane], s1);
        if (!wide_sums) {
          float s2 = transpose_reduce16_(sq, lane);
          if (lane < 16) atomicAdd(&s_stats[prm.J + cbeg + cc + lane], s2);
        }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (prm.stats != nullptr) {
    for (int i = threadIdx.x; i < 2 * prm.J; i += blockDim.x) {
      float s = s_stats[i];
      if (s != 0.f) atomicAdd(&prm.stats[i], (double)s);
    }
  }
  if (prm.colsum != nullptr) {
    for (int i = threadIdx.x; i < prm.J; i += blockDim.x) {
      float s = s_stats[i];
      if (s != 0.f) atomicAdd(&prm.colsum[i], s);
    }
  }
  if (prm.mcast) tc::cluster_sync();          // the peer may still be arriving on this CTA's barriers
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

__global__ void pack3_kernel(const float* __restrict__ w, __half* __restrict__ wp, int dgrad, int J, int K) {
  B2U_PDL_PROLOGUE();
  // fwd: Wp[t][co][ci] = w[t][ci][co];  dgrad: Wp[t][ci][co] = w[8-t][ci][co]
  long long total = 9LL * J * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    long long r = i / K;
    int j = (int)(r % J), t = (int)(r / J);
    float v = dgrad ? w[((long long)(8 - t) * J + j) * K + k] : w[((long long)t * K + k) * J + j];
    wp[i] = __float2half_rn(v);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc3 = nullptr;
bool g_attr3 = false;

int get_enc3() {
  if (g_enc3 != nullptr) return B2U_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B2U_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    b2u_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B2U_ERR_CUDA;
  }
  g_enc3 = (EncodeTiledFn)fn;
  return B2U_OK;
}

CUtensorMapSwizzle swz3(int ks) {
  return ks == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (ks == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

}  // namespace

long long* g_b2u_dbg = nullptr;   // device timeline buffer (b2u_set_option("tc_debug", 1))
int g_b2u_tc_2sm_max_j = 64;  // largest N tile that runs two CTAs per SM
int g_b2u_tc_3sm = 0;         // 1: three CTAs per SM for the low-register variants of thin layers (untested)
int g_b2u_tc_mcast = 0;      // streamed weights: clusters of two CTAs sharing every weight tile by TMA multicast
int g_b2u_tc_max_ctas = 0;   // probe only: cap on the number of CTAs (0 = one or two per SM)
int g_b2u_tc_bgroup = 3;      // streamed weights: 3 = ring slots of one filter row for N <= 128 tiles, 1 = one tile per slot
int g_b2u_tc_halo = 1;       // 0: per-tap loads (conv_tc.cu), 1: halo box, 2: halo box + base_offset, 3: three boxes

int b2u_tc_conv3x3_halo(const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act, void* y,
                        int ldy, int J, double* stats, float* colsum, const void* mask, int ldmask, int mask_act,
                        int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp, void* stream,
                        void* relu_bits_out, const float* post_scale, const float* post_shift) {
  int rc = get_enc3();
  if (rc != B2U_OK) return rc;
  C3Params p{};
  p.N = n; p.H = h; p.W = wd; p.K = K; p.J = J;
  p.KS = K % 64 == 0 ? 64 : (K % 32 == 0 ? 32 : 16);
  p.JT = J <= 256 ? J : 256;
  B2U_REQUIRE(K % 16 == 0 && J % 16 == 0 && J % p.JT == 0, "tc_conv3: unsupported channel counts K=%d J=%d", K, J);
  p.amode = g_b2u_tc_halo == 3 ? 3 : 1;
  p.bo_mode = 0;      // descriptor base offsets are wrong for address-anchored swizzles (probed on B200): unused
  p.y = (__half*)y; p.ldy = ldy; p.bias = bias; p.act = act;
  p.mask = (const __half*)mask; p.ldmask = ldmask; p.mask_act = mask_act; p.accumulate = accumulate; p.stats = stats;
  p.colsum = colsum;
  p.dbg = g_b2u_dbg;
  p.bits_out = (uint8_t*)relu_bits_out;
  p.post_scale = post_scale; p.post_shift = post_shift;
  B2U_REQUIRE((post_scale == nullptr) == (post_shift == nullptr), "tc_conv3: post-activation scale and shift come together");
  if (mask != nullptr && mask_act == B2U_ACT_RELU_BITS) {          // packed 1-bit mask instead of the activation tensor
    p.bits_in = (const uint8_t*)mask;
    p.mask = nullptr;
    mask = nullptr;
  }
  const uint32_t rowb = p.KS * 2;
  const uint32_t rows = p.amode == 3 ? 18 * 8 : 18 * 10;
  p.a_sub = (rows * rowb + 1023) & ~1023u;
  const size_t a_stage = (size_t)p.a_sub * (p.amode == 3 ? 3 : 1);
  const size_t b_tile = ((size_t)p.JT * rowb + 1023) & ~(size_t)1023;
  const int kslabs = K / p.KS;
  const size_t tail = (2 * kMaxSA + 2 * kMaxSB + 5) * 8 + 32 + (size_t)3 * J * 4 + 64;
  const size_t budget = 222 * 1024 - 1024 - tail;
  const size_t wres = 9 * (size_t)kslabs * b_tile;
  p.bres = (p.JT == J && wres + 2 * a_stage <= budget && wres <= 120 * 1024) ? 1 : 0;
  p.bgroup = 1;
  if (p.bres) {
    p.SB = 1;
    p.SA = (int)((budget - wres) / a_stage);
  } else {
    // split the budget: at least 2 A stages, the rest to the weight ring (>= 3 tiles)
    p.SA = 2;
    long long room = (long long)budget - 2 * (long long)a_stage;
    // ring slots of three tiles (one filter row) when the MMAs of one tile are shorter than a commit's issue gap (N <= 128)
    // and two such slots fit: see C3Params::bgroup
    p.bgroup = (g_b2u_tc_bgroup == 3 && p.JT <= 128 && room >= 6 * (long long)b_tile) ? 3 : 1;
    p.SB = (int)(room / ((long long)b_tile * p.bgroup));
    if (p.SB > kMaxSB) {
      p.SB = kMaxSB;
      p.SA = (int)((budget - (size_t)p.SB * p.bgroup * b_tile) / a_stage);
    }
  }
  if (p.SA > kMaxSA) p.SA = kMaxSA;
  B2U_REQUIRE(p.SA >= 2 && p.SB >= 1 && (p.bres || p.SB >= 2), "tc_conv3: tiles do not fit shared memory (K=%d J=%d)", K, J);
  // two CTAs per SM for thin resident-weight layers: cap the A ring so that one CTA stays under ~110 KB.
  // Untuned switches for the next round (defaults = the measured configuration): `tc_2sm_max_j` moves the N-tile limit
  // of the two-CTA mode, `tc_3sm` lets the <= 80-register variants (no statistics, no fp16 mask) run three CTAs per SM.
  bool two_per_sm = false;
  int per_sm = 1;
  const bool lowreg = mask == nullptr && !accumulate && stats == nullptr && colsum == nullptr;
  if (p.bres && p.JT <= g_b2u_tc_2sm_max_j) {
    const bool three = g_b2u_tc_3sm && lowreg && p.JT <= 64;
    const size_t cap = three ? 72 * 1024 : 110 * 1024;
    if (1024 + wres + tail + 2 * a_stage <= cap) {
      int sa2 = (int)((cap - 1024 - wres - tail) / a_stage);
      if (sa2 > kMaxSA) sa2 = kMaxSA;
      if (sa2 >= 2) { p.SA = sa2; two_per_sm = true; per_sm = three ? 3 : 2; }
    }
  }
  p.epi_warps = two_per_sm ? 4 : 8;
  const size_t smem = 1024 + (size_t)p.SA * a_stage + (p.bres ? wres : (size_t)p.SB * p.bgroup * b_tile) + tail;
  B2U_REQUIRE(smem <= 227 * 1024, "tc_conv3: shared memory %zu exceeds 227 KB", smem);

  // packed fp16 weights: the caller's (b2u_pack_weights, once per step for the whole model) or packed here
  if (wp == nullptr) {
    const size_t need = 9 * (size_t)J * K * 2;
    B2U_REQUIRE(ws != nullptr && need <= ws_bytes, "tc_conv3: workspace too small");
    long long total = 9LL * J * K;
    int grid = (int)((total + 255) / 256);
    if (grid > 8 * B2U_NUM_SMS) grid = 8 * B2U_NUM_SMS;
    B2U_LAUNCH(pack3_kernel, grid, 256, 0, stream, w, (__half*)ws, dgrad, J, K);
    wp = ws;
  }
  C3Maps maps;
  {
    cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)wd * ldx * 2, (cuuint64_t)h * wd * ldx * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.KS, (cuuint32_t)(p.amode == 3 ? 8 : 10), 18, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_enc3(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz3(p.KS), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b2u_set_error("tc_conv3: activation tensor map failed (%d)", (int)r); return B2U_ERR_CUDA; }
    cuuint64_t bd[3] = {(cuuint64_t)K, (cuuint64_t)J, 9};
    cuuint64_t bs[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * J * 2};
    cuuint32_t bb[3] = {(cuuint32_t)p.KS, (cuuint32_t)p.JT, 1};
    cuuint32_t be[3] = {1, 1, 1};
    r = g_enc3(&maps.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(wp), bd, bs, bb, be, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swz3(p.KS), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b2u_set_error("tc_conv3: weight tensor map failed (%d)", (int)r); return B2U_ERR_CUDA; }
  }
  if (!g_attr3) {
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2U_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_attr3 = true;
  }
  long long tiles = (long long)n * b2u_cdiv(h, kTH) * b2u_cdiv(wd, kTW) * (J / p.JT);
  B2U_REQUIRE(tiles < (1LL << 31), "tc_conv3: too many tiles");
  // thin layers are bound by the single MMA-issuing thread (~75 issue cycles per tcgen05.mma): run two CTAs
  // per SM (two issuers) when shared memory and TMEM (2*JT columns each) allow it
  int ctas = B2U_NUM_SMS;
  if (two_per_sm && tiles >= 2 * per_sm * B2U_NUM_SMS) ctas = per_sm * B2U_NUM_SMS;
  if (g_b2u_tc_max_ctas > 0 && ctas > g_b2u_tc_max_ctas) ctas = g_b2u_tc_max_ctas;   // probe: fewer CTAs than SMs
  int grid = (int)(tiles < ctas ? tiles : ctas);
  // streamed weights: CTA pairs sharing every weight tile (C3Params::mcast) when there are at least two pixel tiles
  const long long npt = (long long)n * b2u_cdiv(h, kTH) * b2u_cdiv(wd, kTW);
  // (g_b2u_tc_mcast = cluster size 2, 4 or 8: the B300 notes put the L2 de-duplication window of TMA multicast at ~4 CTAs)
  const int cs = (g_b2u_tc_mcast == 2 || g_b2u_tc_mcast == 4 || g_b2u_tc_mcast == 8) ? g_b2u_tc_mcast : 0;
  p.mcast = (cs && !p.bres && !two_per_sm && npt >= cs && p.dbg == nullptr) ? cs : 0;
  if (p.mcast) {
    const long long nsuper = ((npt + cs - 1) / cs) * (J / p.JT);
    long long clusters = ctas / cs;
    int maxc = 0;                      // clusters of this size that can be resident at once (GPC boundaries)
    {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3((unsigned)(cs * clusters)); qc.blockDim = dim3(64 + 32 * p.epi_warps); qc.dynamicSmemBytes = smem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = (unsigned)cs; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&maxc, tc_conv3_kernel<0>, &qc) != cudaSuccess) { cudaGetLastError(); maxc = 0; }
    }
    if (maxc > 0 && clusters > maxc) clusters = maxc;
    if (clusters > nsuper) clusters = nsuper;
    if (clusters < 1) p.mcast = 0; else grid = (int)(cs * clusters);
  }
  const int flags = ((mask != nullptr || accumulate) ? kF_MASKACC : 0) | ((stats != nullptr || colsum != nullptr) ? kF_SUMS : 0) |
                    (p.bits_out != nullptr ? kF_BITS_OUT : 0) | (p.bits_in != nullptr ? kF_BITS_IN : 0) |
                    (post_scale != nullptr ? kF_POST : 0);
  const int nthr = 64 + 32 * p.epi_warps;
  if (p.mcast) {
    switch (flags) {
      case 0: B2U_LAUNCH_CLUSTER(tc_conv3_kernel<0>, grid, nthr, smem, stream, p.mcast, maps, p); break;
      case 1: B2U_LAUNCH_CLUSTER(tc_conv3_kernel<1>, grid, nthr, smem, stream, p.mcast, maps, p); break;
      case 2: B2U_LAUNCH_CLUSTER(tc_conv3_kernel<2>, grid, nthr, smem, stream, p.mcast, maps, p); break;
      case 3: B2U_LAUNCH_CLUSTER(tc_conv3_kernel<3>, grid, nthr, smem, stream, p.mcast, maps, p); break;
      case 4: B2U_LAUNCH_CLUSTER(tc_conv3_kernel<4>, grid, nthr, smem, stream, p.mcast, maps, p); break;
      case 6: B2U_LAUNCH_CLUSTER(tc_conv3_kernel<6>, grid, nthr, smem, stream, p.mcast, maps, p); break;
      case 8: B2U_LAUNCH_CLUSTER(tc_conv3_kernel<8>, grid, nthr, smem, stream, p.mcast, maps, p); break;
      case 10: B2U_LAUNCH_CLUSTER(tc_conv3_kernel<10>, grid, nthr, smem, stream, p.mcast, maps, p); break;
      case 16: B2U_LAUNCH_CLUSTER(tc_conv3_kernel<16>, grid, nthr, smem, stream, p.mcast, maps, p); break;
      default:
        b2u_set_error("tc_conv3: unsupported feature combination %d", flags);
        return B2U_ERR_ARG;
    }
    return B2U_OK;
  }
  switch (flags) {
    case 0: B2U_LAUNCH(tc_conv3_kernel<0>, grid, nthr, smem, stream, maps, p); break;
    case 1: B2U_LAUNCH(tc_conv3_kernel<1>, grid, nthr, smem, stream, maps, p); break;
    case 2: B2U_LAUNCH(tc_conv3_kernel<2>, grid, nthr, smem, stream, maps, p); break;
    case 3: B2U_LAUNCH(tc_conv3_kernel<3>, grid, nthr, smem, stream, maps, p); break;
    case 4: B2U_LAUNCH(tc_conv3_kernel<4>, grid, nthr, smem, stream, maps, p); break;      // forward + bit mask
    case 6: B2U_LAUNCH(tc_conv3_kernel<6>, grid, nthr, smem, stream, maps, p); break;      // ... + statistics
    case 8: B2U_LAUNCH(tc_conv3_kernel<8>, grid, nthr, smem, stream, maps, p); break;      // data gradient, 1-bit mask
    case 10: B2U_LAUNCH(tc_conv3_kernel<10>, grid, nthr, smem, stream, maps, p); break;    // ... + column sums
    case 16: B2U_LAUNCH(tc_conv3_kernel<16>, grid, nthr, smem, stream, maps, p); break;    // inference forward + BN affine
    default:
      b2u_set_error("tc_conv3: unsupported feature combination %d (1-bit masks do not combine with accumulate / fp16 masks, the post-activation affine only with a plain forward)", flags);
      return B2U_ERR_ARG;
  }
  return B2U_OK;
}
