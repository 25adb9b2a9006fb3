// Exact-mode (fp32 accumulate on CUDA cores) convolution kernels: 3x3 'same' conv forward / dgrad /
// wgrad and the 2x2 stride-2 transposed conv.  These serve (a) dt == B2U_F32, the bit-faithful
// parity mode, (b) the layers tensor cores cannot help with (Cin == 1 first conv: K = 9), and
// (c) the cross-check of the tcgen05 kernels in conv_tc.cu.  Shared-memory tiled, NHWC, coalesced.
#include "common.cuh"
#include "launch.cuh"

namespace {

// ==========================================================================================
// 3x3 conv, stride 1, pad 1.  One kernel serves forward and dgrad through a weight accessor:
//   fwd   : Wacc(tap,k,j) = w[(tap*K + k)*J + j]            (HWIO, K = Cin,  J = Cout)
//   dgrad : Wacc(tap,k,j) = w[((8-tap)*J + j)*K + k]        (HWIO, K = Cout, J = Cin, taps rotated)
// Block = 16x16 output pixels x TCO output channels, 256 threads, each 8 pixels (row segment) x 4 ch
// (TCO=32) or 16 pixels x 4 ch (TCO=64).
// ==========================================================================================
constexpr int TH = 16, TW = 16, KC = 8;
constexpr int XS_W = TW + 2 + 1;   // padded row length of the halo tile in smem

template <typename T, int TCO>
__global__ void __launch_bounds__(256) conv3x3_direct_kernel(
    const T* __restrict__ x, int ldx, int K, const float* __restrict__ w, int dgrad,
    const float* __restrict__ bias, int act, T* __restrict__ y, int ldy, int J, double* __restrict__ stats,
    const T* __restrict__ mask, int ldmask, int mask_act, int accumulate, int N, int H, int W) {
  B2U_PDL_PROLOGUE();
  constexpr int NCG = TCO / 4;             // channel groups of 4
  constexpr int NPG = 256 / NCG;           // pixel groups
  constexpr int PX = (TH * TW) / NPG;      // pixels per thread: 8 (TCO=32) or 16 (TCO=64)
  __shared__ float xs[KC][TH + 2][XS_W];
  __shared__ __align__(16) float ws[9][KC][TCO];
  __shared__ double sstat[2 * TCO];

  const int tiles_w = (W + TW - 1) / TW;
  const int tile_h0 = (blockIdx.x / tiles_w) * TH, tile_w0 = (blockIdx.x % tiles_w) * TW;
  const int n = blockIdx.y;
  const int j0 = blockIdx.z * TCO;
  const int tid = threadIdx.x;
  const int cgid = tid % NCG, pg = tid / NCG;
  // thread's pixels: row r, columns c0 .. c0+PX-1 (PX=8: two segments per row; PX=16: a full row)
  const int r = (pg * PX) / TW, c0 = (pg * PX) % TW;

  float acc[PX][4];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;

  if (stats != nullptr) {
    for (int i = tid; i < 2 * TCO; i += 256) sstat[i] = 0.0;
  }

  for (int k0 = 0; k0 < K; k0 += KC) {
    __syncthreads();
    // ---- halo tile: (TH+2) x (TW+2) pixels x KC channels, zero outside the image -------------
    for (int i = tid; i < (TH + 2) * (TW + 2); i += 256) {
      int rr = i / (TW + 2), cc = i % (TW + 2);
      int hh = tile_h0 + rr - 1, wwp = tile_w0 + cc - 1;
      float v[KC];
#pragma unroll
      for (int q = 0; q < KC; ++q) v[q] = 0.f;
      if (hh >= 0 && hh < H && wwp >= 0 && wwp < W) {
        const T* src = x + (((long long)n * H + hh) * W + wwp) * ldx + k0;
        if ((K & 7) == 0) {
          load8<T>(src, v);
        } else {
          for (int q = 0; q < KC; ++q)
            if (k0 + q < K) v[q] = ldf<T>(src + q);
        }
      }
#pragma unroll
      for (int q = 0; q < KC; ++q) xs[q][rr][cc] = v[q];
    }
    // ---- weight tile [9][KC][TCO] ---------------------------------------------------------------
    for (int i = tid; i < 9 * KC * TCO; i += 256) {
      int jj = i % TCO, kk = (i / TCO) % KC, tap = i / (TCO * KC);
      int k = k0 + kk, j = j0 + jj;
      float v = 0.f;
      if (k < K && j < J) {
        v = dgrad ? __ldg(w + ((long long)(8 - tap) * J + j) * K + k) : __ldg(w + ((long long)tap * K + k) * J + j);
      }
      ws[tap][kk][jj] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int kk = 0; kk < KC; ++kk) {
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        float xv[PX + 2];
#pragma unroll
        for (int q = 0; q < PX + 2; ++q) xv[q] = xs[kk][r + dh][c0 + q];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          float4 wv = *reinterpret_cast<const float4*>(&ws[dh * 3 + dw][kk][cgid * 4]);
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            acc[p][0] = fmaf(xv[p + dw], wv.x, acc[p][0]);
            acc[p][1] = fmaf(xv[p + dw], wv.y, acc[p][1]);
            acc[p][2] = fmaf(xv[p + dw], wv.z, acc[p][2]);
            acc[p][3] = fmaf(xv[p + dw], wv.w, acc[p][3]);
          }
        }
      }
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------
  const int jb = j0 + cgid * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (jb + q < J) bv[q] = __ldg(bias + jb + q);
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const int hh = tile_h0 + r;
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    int wwp = tile_w0 + c0 + p;
    if (hh < H && wwp < W) {
      long long pix = ((long long)n * H + hh) * W + wwp;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (jb + q < J) {
          float v = act_fwd(acc[p][q] + bv[q], act);
          if (mask != nullptr) v *= act_bwd_from_y(ldf<T>(mask + pix * ldmask + jb + q), mask_act);
          T* dst = y + pix * ldy + jb + q;
          if (accumulate) v += ldf<T>(dst);
          stf<T>(dst, v);
          if (stats != nullptr) {
            float vr = ldf<T>(dst);   // statistics of the value as stored (matters for fp16 storage)
            s1[q] += vr;
            s2[q] += vr * vr;
          }
        }
      }
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      atomicAdd(&sstat[cgid * 4 + q], (double)s1[q]);
      atomicAdd(&sstat[TCO + cgid * 4 + q], (double)s2[q]);
    }
    __syncthreads();
    for (int i = tid; i < TCO; i += 256) {
      if (j0 + i < J) {
        atomicAdd(&stats[j0 + i], sstat[i]);
        atomicAdd(&stats[J + j0 + i], sstat[TCO + i]);
      }
    }
  }
}

// ==========================================================================================
// Cin == 1 first layer (T1H:859 Conv2D(32,(3,3)) on the (H,W,1) input; T2:748 with 16 filters).
// K = 9: no tensor-core shape, pure HBM streaming (writes COUT x what it reads).  One thread per pixel,
// 64-byte (fp16) contiguous stores; the backward accumulates the 9 x COUT weight gradient in registers.
// ==========================================================================================
// Thread = (pixel lane, 8-channel group): the 9 x 8 weights of the group live in registers, a warp's 16-byte
// stores cover 8 pixels x COUT channels = one contiguous 512-byte (COUT = 32) span, and the x halo of the NEXT
// tile is fetched into registers while the current one is computed (persistent blocks, 2 per SM).
constexpr int C1_TH = 8, C1_TW = 32;                        // pixel tile of the Cin == 1 kernels
constexpr int C1_HALO = (C1_TH + 2) * (C1_TW + 2);          // 612 halo elements
constexpr int C1_XR = (C1_HALO + 255) / 256;                // halo elements per thread

template <typename T>
__device__ __forceinline__ void c1_load_halo(const T* __restrict__ x, int ldx, int H, int W, int tile, int tiles_w,
                                             int tiles_h, float xr[C1_XR]) {
  const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
  const int h0 = th * C1_TH - 1, w0 = tw * C1_TW - 1;
#pragma unroll
  for (int k = 0; k < C1_XR; ++k) {
    const int i = threadIdx.x + k * 256;
    const int rr = i / (C1_TW + 2), cc = i % (C1_TW + 2), hh = h0 + rr, wwp = w0 + cc;
    float v = 0.f;
    if (i < C1_HALO && hh >= 0 && hh < H && wwp >= 0 && wwp < W) v = ldf<T>(x + (((long long)n * H + hh) * W + wwp) * ldx);
    xr[k] = v;
  }
}

template <typename T, int COUT>
__global__ void __launch_bounds__(256, 2) conv3x3_c1_fwd_kernel(const T* __restrict__ x, int ldx,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ bias, int act,
                                                                T* __restrict__ y, int ldy, int N, int H, int W,
                                                                uint8_t* __restrict__ bits) {
  B2U_PDL_PROLOGUE();
  constexpr int CG = COUT / 8, LANES = 256 / CG;
  // a stored value is > 0 iff the fp32 value survives the rounding to the storage type
  const float pos = sizeof(T) == 2 ? 2.98023223876953125e-08f : 0.f;
  __shared__ float xs[C1_HALO + 2];
  const int g = threadIdx.x % CG, lane = threadIdx.x / CG;
  float wr[9][8], br[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    br[k] = bias ? bias[g * 8 + k] : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][k] = w[t * COUT + g * 8 + k];
  }
  const int tiles_w = (W + C1_TW - 1) / C1_TW, tiles_h = (H + C1_TH - 1) / C1_TH;
  const int ntiles = N * tiles_h * tiles_w;
  float xr[C1_XR];
  if ((int)blockIdx.x < ntiles) c1_load_halo<T>(x, ldx, H, W, blockIdx.x, tiles_w, tiles_h, xr);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int h0 = th * C1_TH, w0 = tw * C1_TW;
    __syncthreads();                                         // the previous tile's readers are done
#pragma unroll
    for (int k = 0; k < C1_XR; ++k)
      if (threadIdx.x + k * 256 < C1_HALO) xs[threadIdx.x + k * 256] = xr[k];
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) c1_load_halo<T>(x, ldx, H, W, tile + gridDim.x, tiles_w, tiles_h, xr);
    // two horizontally adjacent pixels per thread and trip: the 3 x 4 input window is read once for both (12 shared-
    // memory loads instead of 18) and the index arithmetic is shared -- the kernel is instruction-issue bound (ncu,
    // round 2: 213 instructions per pixel and channel group for its 72 FMAs, 63 % issue utilisation at 16 warps per SM)
#pragma unroll 2
    for (int pp = lane; pp < C1_TH * C1_TW / 2; pp += LANES) {
      const int r = pp / (C1_TW / 2), c = (pp % (C1_TW / 2)) * 2;
      if (h0 + r >= H || w0 + c >= W) continue;
      float o0[8], o1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { o0[k] = br[k]; o1[k] = br[k]; }
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        float xw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) xw[q] = xs[(r + dh) * (C1_TW + 2) + c + q];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            o0[k] = fmaf(xw[dw], wr[dh * 3 + dw][k], o0[k]);
            o1[k] = fmaf(xw[dw + 1], wr[dh * 3 + dw][k], o1[k]);
          }
      }
      if (act == B2U_ACT_RELU) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { o0[k] = fmaxf(o0[k], 0.f); o1[k] = fmaxf(o1[k], 0.f); }
      } else if (act != B2U_ACT_NONE) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { o0[k] = act_fwd(o0[k], act); o1[k] = act_fwd(o1[k], act); }
      }
      const long long pix = ((long long)n * H + h0 + r) * W + w0 + c;
      const bool second = w0 + c + 1 < W;
      store8<T>(y + pix * ldy + g * 8, o0);
      if (second) store8<T>(y + (pix + 1) * ldy + g * 8, o1);
      if (bits != nullptr) {               // packed 1-bit ReLU mask (bit pix*COUT + channel): one byte per thread and pixel
        unsigned b0 = 0u, b1 = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) { b0 |= (o0[k] > pos ? 1u : 0u) << k; b1 |= (o1[k] > pos ? 1u : 0u) << k; }
        bits[pix * CG + g] = (uint8_t)b0;
        if (second) bits[(pix + 1) * CG + g] = (uint8_t)b1;
      }
    }
  }
}

template <typename T, int COUT>
__global__ void __launch_bounds__(256, 2) conv3x3_c1_wgrad_kernel(const T* __restrict__ x, int ldx,
                                                                  const T* __restrict__ dy, int lddy,
                                                                  float* __restrict__ dw, float* __restrict__ db, int N,
                                                                  int H, int W) {
  B2U_PDL_PROLOGUE();
  // block = 8 x 32 pixel tiles (grid-stride); thread = (pixel lane, 8-channel group).  The loads of the NEXT tile
  // (the thread's share of the x halo, its pixels' 16-byte dy pieces) are issued before the current tile is
  // reduced, so a block keeps ~17 KB in flight all the time; 9 taps x 8 channels (+ bias) of the gradient
  // accumulate in registers over all tiles of the block and are reduced once at the end.
  constexpr int CG = COUT / 8;
  constexpr int LANES = 256 / CG;
  constexpr int PPT = C1_TH * C1_TW / LANES;                 // pixels per thread and tile (4 at COUT = 32)
  __shared__ float xs[C1_HALO + 2];
  __shared__ float sacc[10 * COUT];
  for (int i = threadIdx.x; i < 10 * COUT; i += 256) sacc[i] = 0.f;
  const int g = threadIdx.x % CG, lane = threadIdx.x / CG;
  float acc[9][8], accb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t][k] = 0.f;
  }
  const int tiles_w = (W + C1_TW - 1) / C1_TW, tiles_h = (H + C1_TH - 1) / C1_TH;
  const int ntiles = N * tiles_h * tiles_w;
  float xr[C1_XR];
  constexpr int RW = sizeof(T) == 2 ? 1 : 2;                 // 16-byte words per 8 channels
  uint4 dn[PPT][RW];                                         // next tile's dy (prefetched, still packed)
  auto load_dy = [&](int tile) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int h0 = th * C1_TH, w0 = tw * C1_TW;
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int pp = lane + q * LANES;
      const int r = pp / C1_TW, c = pp % C1_TW;
#pragma unroll
      for (int i = 0; i < RW; ++i) dn[q][i] = make_uint4(0u, 0u, 0u, 0u);
      if (h0 + r < H && w0 + c < W) {
        const uint4* src = reinterpret_cast<const uint4*>(dy + (((long long)n * H + h0 + r) * W + w0 + c) * lddy + g * 8);
#pragma unroll
        for (int i = 0; i < RW; ++i) dn[q][i] = src[i];
      }
    }
  };
  if ((int)blockIdx.x < ntiles) {
    c1_load_halo<T>(x, ldx, H, W, blockIdx.x, tiles_w, tiles_h, xr);
    load_dy(blockIdx.x);
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();                                         // the previous tile's readers are done
#pragma unroll
    for (int k = 0; k < C1_XR; ++k)
      if (threadIdx.x + k * 256 < C1_HALO) xs[threadIdx.x + k * 256] = xr[k];
    uint4 dc[PPT][RW];
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
      for (int i = 0; i < RW; ++i) dc[q][i] = dn[q][i];
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) {                    // next tile's loads fly while this one is reduced
      c1_load_halo<T>(x, ldx, H, W, tile + gridDim.x, tiles_w, tiles_h, xr);
      load_dy(tile + gridDim.x);
    }
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int pp = lane + q * LANES;
      const int r = pp / C1_TW, c = pp % C1_TW;
      float d[8];
      load8<T>(reinterpret_cast<const T*>(&dc[q][0]), d);    // unpack from registers
#pragma unroll
      for (int k = 0; k < 8; ++k) accb[k] += d[k];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float xv = xs[(r + t / 3) * (C1_TW + 2) + c + t % 3];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(xv, d[k], acc[t][k]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float vb = accb[k];
    for (int o = 16; o >= CG; o >>= 1) vb += __shfl_xor_sync(0xffffffffu, vb, o);
    if ((threadIdx.x & 31) < CG) atomicAdd(&sacc[9 * COUT + g * 8 + k], vb);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v = acc[t][k];
      for (int o = 16; o >= CG; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) < CG) atomicAdd(&sacc[t * COUT + g * 8 + k], v);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * COUT; i += 256) atomicAdd(dw + i, sacc[i]);
  if (db != nullptr)
    for (int i = threadIdx.x; i < COUT; i += 256) atomicAdd(db + i, sacc[9 * COUT + i]);
}

// ==========================================================================================
// 3x3 wgrad: dw[tap][ci][co] += sum_p x[p+tap][ci] * dy[p][co];  db[co] += sum_p dy[p][co]
// Block owns (ci chunk of 8) x (co tile of 64) x all 9 taps and loops over its share of 8x16 pixel
// tiles, then flushes once with atomics.  128 threads: 32 co-pairs x 4 ci-pairs.
// ==========================================================================================
constexpr int WG_TH = 8, WG_TW = 16, WG_CO = 64, WG_CI = 8;

template <typename T>
__global__ void __launch_bounds__(128) conv3x3_wgrad_direct_kernel(const T* __restrict__ x, int ldx, int Cin,
                                                                   const T* __restrict__ dy, int lddy, int Cout,
                                                                   float* __restrict__ dw, float* __restrict__ db,
                                                                   int N, int H, int W, int nsplit) {
  B2U_PDL_PROLOGUE();
  __shared__ __align__(16) float xs[WG_TH + 2][WG_TW + 2][WG_CI];
  __shared__ __align__(16) float ds[WG_TH][WG_TW][WG_CO];
  const int ci0 = blockIdx.y * WG_CI, co0 = blockIdx.z * WG_CO;
  const int tid = threadIdx.x;
  const int cop = tid % 32, cip = tid / 32;    // co pair, ci pair
  float acc[9][2][2];
#pragma unroll
  for (int t = 0; t < 9; ++t) { acc[t][0][0] = acc[t][0][1] = acc[t][1][0] = acc[t][1][1] = 0.f; }
  float accb[2] = {0.f, 0.f};
  const int tiles_w = (W + WG_TW - 1) / WG_TW, tiles_h = (H + WG_TH - 1) / WG_TH;
  const long long ntiles = (long long)N * tiles_h * tiles_w;
  for (long long tile = blockIdx.x; tile < ntiles; tile += nsplit) {
    int tw = (int)(tile % tiles_w);
    long long tt = tile / tiles_w;
    int th = (int)(tt % tiles_h);
    int n = (int)(tt / tiles_h);
    int h0 = th * WG_TH, w0 = tw * WG_TW;
    __syncthreads();
    for (int i = tid; i < (WG_TH + 2) * (WG_TW + 2); i += 128) {
      int rr = i / (WG_TW + 2), cc = i % (WG_TW + 2);
      int hh = h0 + rr - 1, wwp = w0 + cc - 1;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = 0.f;
      if (hh >= 0 && hh < H && wwp >= 0 && wwp < W) {
        const T* src = x + (((long long)n * H + hh) * W + wwp) * ldx + ci0;
        if ((Cin & 7) == 0) load8<T>(src, v);
        else
          for (int q = 0; q < 8; ++q)
            if (ci0 + q < Cin) v[q] = ldf<T>(src + q);
      }
      *reinterpret_cast<float4*>(&xs[rr][cc][0]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&xs[rr][cc][4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    for (int i = tid; i < WG_TH * WG_TW * (WG_CO / 8); i += 128) {
      int g = i % (WG_CO / 8), pp = i / (WG_CO / 8);
      int rr = pp / WG_TW, cc = pp % WG_TW;
      int hh = h0 + rr, wwp = w0 + cc;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = 0.f;
      if (hh < H && wwp < W && co0 + g * 8 < Cout) {
        const T* src = dy + (((long long)n * H + hh) * W + wwp) * lddy + co0 + g * 8;
        if ((Cout & 7) == 0) load8<T>(src, v);
        else
          for (int q = 0; q < 8; ++q)
            if (co0 + g * 8 + q < Cout) v[q] = ldf<T>(src + q);
      }
      *reinterpret_cast<float4*>(&ds[rr][cc][g * 8]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&ds[rr][cc][g * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    for (int rr = 0; rr < WG_TH; ++rr) {
#pragma unroll 4
      for (int cc = 0; cc < WG_TW; ++cc) {
        float2 d = *reinterpret_cast<const float2*>(&ds[rr][cc][cop * 2]);
        accb[0] += d.x;
        accb[1] += d.y;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float2 xv = *reinterpret_cast<const float2*>(&xs[rr + t / 3][cc + t % 3][cip * 2]);
          acc[t][0][0] = fmaf(xv.x, d.x, acc[t][0][0]);
          acc[t][0][1] = fmaf(xv.x, d.y, acc[t][0][1]);
          acc[t][1][0] = fmaf(xv.y, d.x, acc[t][1][0]);
          acc[t][1][1] = fmaf(xv.y, d.y, acc[t][1][1]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        int ci = ci0 + cip * 2 + a, co = co0 + cop * 2 + b;
        if (ci < Cin && co < Cout) atomicAdd(dw + ((long long)t * Cin + ci) * Cout + co, acc[t][a][b]);
      }
  if (db != nullptr && blockIdx.y == 0 && cip == 0) {
#pragma unroll
    for (int b = 0; b < 2; ++b)
      if (co0 + cop * 2 + b < Cout) atomicAdd(db + co0 + cop * 2 + b, accb[b]);
  }
}

// ==========================================================================================
// Generic 64x64x16 smem GEMM with functor loaders, used by the three transposed-conv passes.
//   C[m][n] = sum_k A(m,k) * B(k,n)
// ==========================================================================================
template <typename LA, typename LB, typename EP>
__global__ void __launch_bounds__(256) gemm64_kernel(LA la, LB lb, EP ep, long long M, int Nn, long long Kk,
                                                     int ksplit) {
  B2U_PDL_PROLOGUE();
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  // split-K: this block reduces k in [kbeg, kend)
  long long kchunk = ((Kk + ksplit - 1) / ksplit + 15) / 16 * 16;
  long long kbeg = (long long)blockIdx.z * kchunk, kend = kbeg + kchunk < Kk ? kbeg + kchunk : Kk;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long k0 = kbeg; k0 < kend; k0 += 16) {
    __syncthreads();
    for (int i = tid; i < 64 * 16; i += 256) {
      int kk = i % 16, mm = i / 16;     // k fastest: A is K-contiguous for most callers
      long long m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < kend) ? la(m, k) : 0.f;
    }
    for (int i = tid; i < 64 * 16; i += 256) {
      int nn = i % 64, kk = i / 64;
      long long k = k0 + kk;
      int nidx = n0 + nn;
      Bs[kk][nn] = (nidx < Nn && k < kend) ? lb(k, nidx) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long m = m0 + ty * 4 + i;
      int nidx = n0 + tx * 4 + j;
      if (m < M && nidx < Nn) ep(m, nidx, acc[i][j]);
    }
}

// ---- transposed conv functors -------------------------------------------------------------------
// pixel index helpers: p = (n*H + i)*W + j over the INPUT grid; q = (a*2+b)*Cout + co
template <typename T>
struct CtFwdA {   // A(p, ci) = x[p][ci]
  const T* x; int ldx;
  __device__ float operator()(long long p, long long ci) const { return ldf<T>(x + p * ldx + ci); }
};
struct CtFwdB {   // B(ci, q) = w[q*Cin + ci]
  const float* w; int Cin;
  __device__ float operator()(long long ci, int q) const { return __ldg(w + (long long)q * Cin + ci); }
};
template <typename T>
struct CtFwdEp {
  T* y; int ldy; int Cout; int H, W; const float* bias;
  __device__ void operator()(long long p, int q, float v) const {
    int ab = q / Cout, co = q - ab * Cout;
    int j = (int)(p % W);
    long long t = p / W;
    int i = (int)(t % H);
    long long n = t / H;
    long long op = ((n * 2 * H + 2 * i + (ab >> 1)) * (2LL * W) + 2 * j + (ab & 1));
    stf<T>(y + op * ldy + co, v + __ldg(bias + co));
  }
};
template <typename T>
struct CtDgradA {   // A(p, q) = dy[(n,2i+a,2j+b)][co]
  const T* dy; int lddy; int Cout; int H, W;
  __device__ float operator()(long long p, long long q) const {
    int ab = (int)(q / Cout), co = (int)(q - (long long)ab * Cout);
    int j = (int)(p % W);
    long long t = p / W;
    int i = (int)(t % H);
    long long n = t / H;
    long long op = ((n * 2 * H + 2 * i + (ab >> 1)) * (2LL * W) + 2 * j + (ab & 1));
    return ldf<T>(dy + op * lddy + co);
  }
};
struct CtDgradB {   // B(q, ci) = w[q*Cin + ci]
  const float* w; int Cin;
  __device__ float operator()(long long q, int ci) const { return __ldg(w + q * Cin + ci); }
};
template <typename T>
struct CtDgradEp {
  T* dx; int lddx; const T* mask; int ldmask; int mask_act; int accumulate;
  __device__ void operator()(long long p, int ci, float v) const {
    if (mask != nullptr) v *= act_bwd_from_y(ldf<T>(mask + p * ldmask + ci), mask_act);
    if (accumulate) v += ldf<T>(dx + p * lddx + ci);
    stf<T>(dx + p * lddx + ci, v);
  }
};
template <typename T>
struct CtWgradA {   // A(q, p) = dy[(n,2i+a,2j+b)][co]
  const T* dy; int lddy; int Cout; int H, W;
  __device__ float operator()(long long q, long long p) const {
    int ab = (int)(q / Cout), co = (int)(q - (long long)ab * Cout);
    int j = (int)(p % W);
    long long t = p / W;
    int i = (int)(t % H);
    long long n = t / H;
    long long op = ((n * 2 * H + 2 * i + (ab >> 1)) * (2LL * W) + 2 * j + (ab & 1));
    return ldf<T>(dy + op * lddy + co);
  }
};
template <typename T>
struct CtWgradB {   // B(p, ci) = x[p][ci]
  const T* x; int ldx;
  __device__ float operator()(long long p, int ci) const { return ldf<T>(x + p * ldx + ci); }
};
struct CtWgradEp {
  float* dw; int Cin;
  __device__ void operator()(long long q, int ci, float v) const { atomicAdd(dw + q * Cin + ci, v); }
};

// db[co] += sum over all output pixels of dy[.][co]
template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(const T* __restrict__ dy, int lddy, int C, long long npix,
                                                          float* __restrict__ db) {
  B2U_PDL_PROLOGUE();
  extern __shared__ float sacc[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cg = C >> 3;
  const int lanes = 256 / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  {
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    if (lane < lanes) {
      for (long long p = (long long)blockIdx.x * lanes + lane; p < npix; p += (long long)gridDim.x * lanes) {
        float v[8];
        load8<T>(dy + p * lddy + g * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += v[i];
      }
    }
    if (lane < lanes || (cg & (cg - 1)) == 0) group_add8(sacc, s, cg, g);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&db[i], sacc[i]);
}

// ==========================================================================================
// 2 <= Cin <= 4 first layer (Task-2 classifier on 224 x 224 x 3 slices, T2:748: Conv2D(16,(3,3)) -- BASELINE configs[4]).
// K = 9 * Cin = 27 is no tensor-core shape either; the generic direct kernel spent 0.46 ms of the classifier's 0.83 ms
// forward here (B200, batch 64).  Same scheme as the Cin == 1 kernel: thread = (pixel pair, 2-channel group), the
// 9 * Cin * 2 weights of the group live in registers, the halo tile (all Cin channels, NHWC = contiguous) is staged in
// shared memory, a thread computes two horizontally adjacent pixels so that the 3 x 4 input window is read once for both.
// ==========================================================================================
template <typename T> __device__ __forceinline__ void store4(T* p, const float v[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, const float v[4]) {
  const __half2 lo = __floats2half2_rn(v[0], v[1]), hi = __floats2half2_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
}

template <typename T> __device__ __forceinline__ void store2(T* p, const float v[2]);
template <> __device__ __forceinline__ void store2<float>(float* p, const float v[2]) {
  *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}
template <> __device__ __forceinline__ void store2<__half>(__half* p, const float v[2]) {
  *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v[0], v[1]);
}

template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(256, 2) conv3x3_smallcin_fwd_kernel(const T* __restrict__ x, int ldx,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ bias, int act,
                                                                   T* __restrict__ y, int ldy, int N, int H, int W) {
  B2U_PDL_PROLOGUE();
  // 2 channels per thread: 54 weight registers, two blocks per SM (with 4 channels the 108 weight registers allowed one
  // block of 8 warps per SM and the kernel was latency-bound: 0.27 ms for the classifier's first conv at batch 64)
  constexpr int CPT = 2, CG = COUT / CPT, LANES = 256 / CG;
  constexpr int HW2 = C1_TW + 2, HALO = (C1_TH + 2) * HW2;
  __shared__ float xs[HALO * CIN];
  const int g = threadIdx.x % CG, lane = threadIdx.x / CG;
  float wr[9 * CIN][CPT], br[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    br[k] = bias ? bias[g * CPT + k] : 0.f;
#pragma unroll
    for (int t = 0; t < 9 * CIN; ++t) wr[t][k] = w[t * COUT + g * CPT + k];       // HWIO: [tap][ci][co]
  }
  const int tiles_w = (W + C1_TW - 1) / C1_TW, tiles_h = (H + C1_TH - 1) / C1_TH;
  const int ntiles = N * tiles_h * tiles_w;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int h0 = th * C1_TH, w0 = tw * C1_TW;
    __syncthreads();                                         // the previous tile's readers are done
    for (int i = threadIdx.x; i < HALO * CIN; i += 256) {
      const int pix = i / CIN, ci = i % CIN;
      const int hh = h0 - 1 + pix / HW2, ww = w0 - 1 + pix % HW2;
      float v = 0.f;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = ldf<T>(x + (((long long)n * H + hh) * W + ww) * ldx + ci);
      xs[i] = v;
    }
    __syncthreads();
    for (int pp = lane; pp < C1_TH * C1_TW / 2; pp += LANES) {
      const int r = pp / (C1_TW / 2), c = (pp % (C1_TW / 2)) * 2;
      float a0[CPT], a1[CPT];
#pragma unroll
      for (int k = 0; k < CPT; ++k) { a0[k] = br[k]; a1[k] = br[k]; }
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        float xv[4][CIN];                                    // the row's four columns c-1 .. c+2 (halo coordinates c .. c+3)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) xv[q][ci] = xs[((r + dh) * HW2 + c + q) * CIN + ci];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw)
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
              a0[k] = fmaf(xv[dw][ci], wr[(dh * 3 + dw) * CIN + ci][k], a0[k]);
              a1[k] = fmaf(xv[dw + 1][ci], wr[(dh * 3 + dw) * CIN + ci][k], a1[k]);
            }
      }
#pragma unroll
      for (int k = 0; k < CPT; ++k) { a0[k] = act_fwd(a0[k], act); a1[k] = act_fwd(a1[k], act); }
      if (h0 + r < H) {
        T* yp = y + (((long long)n * H + h0 + r) * W + w0 + c) * ldy + g * CPT;
        if (w0 + c < W) store2<T>(yp, a0);                   // a pixel's COUT channels = one contiguous 32 / 64-byte span
        if (w0 + c + 1 < W) store2<T>(yp + ldy, a1);
      }
    }
  }
}

}  // namespace

// ==========================================================================================
// host-side launchers (internal linkage names used by api.cu dispatch)
// ==========================================================================================
#define DISPATCH_T(dt, ...)                                              \
  if ((dt) == B2U_F32) { using T = float; __VA_ARGS__; }                 \
  else if ((dt) == B2U_F16) { using T = __half; __VA_ARGS__; }           \
  else { b2u_set_error("bad dtype %d", (int)(dt)); return B2U_ERR_ARG; }

int b2u_direct_conv3x3(int dt, const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act,
                       void* y, int ldy, int J, double* stats, const void* mask, int ldmask, int mask_act,
                       int accumulate, int n, int h, int wd, void* stream, void* relu_bits, int* bits_done) {
  if (bits_done != nullptr) *bits_done = 0;
  B2U_REQUIRE(n > 0 && h > 0 && wd > 0 && K > 0 && J > 0, "conv3x3: empty shape");
  B2U_REQUIRE((K % 8 != 0) || (ldx % 8 == 0), "conv3x3: ldx must be a multiple of 8 when Cin %% 8 == 0");
  int tiles = b2u_cdiv(h, TH) * b2u_cdiv(wd, TW);
  if (K == 1 && !dgrad && stats == nullptr && mask == nullptr && !accumulate && ldy % 8 == 0 && (J == 32 || J == 16)) {
    long long gl1 = (long long)n * b2u_cdiv(h, C1_TH) * b2u_cdiv(wd, C1_TW);
    int grid1 = (int)(gl1 < 2 * B2U_NUM_SMS ? gl1 : 2 * B2U_NUM_SMS);
    if (J == 32) { DISPATCH_T(dt, B2U_LAUNCH((conv3x3_c1_fwd_kernel<T, 32>), grid1, 256, 0, stream, (const T*)x, ldx, w, bias, act, (T*)y, ldy, n, h, wd, (uint8_t*)relu_bits)); }
    else { DISPATCH_T(dt, B2U_LAUNCH((conv3x3_c1_fwd_kernel<T, 16>), grid1, 256, 0, stream, (const T*)x, ldx, w, bias, act, (T*)y, ldy, n, h, wd, (uint8_t*)relu_bits)); }
    if (bits_done != nullptr && relu_bits != nullptr) *bits_done = 1;
    return B2U_OK;
  }
  if (K == 3 && !dgrad && stats == nullptr && mask == nullptr && !accumulate && (J == 16 || J == 32) && ldy % 4 == 0 &&
      ((uintptr_t)y & 15) == 0) {
    long long gl = (long long)n * b2u_cdiv(h, C1_TH) * b2u_cdiv(wd, C1_TW);
    int grid3 = (int)(gl < 2 * B2U_NUM_SMS ? gl : 2 * B2U_NUM_SMS);       // two persistent blocks per SM
    if (J == 16) { DISPATCH_T(dt, B2U_LAUNCH((conv3x3_smallcin_fwd_kernel<T, 3, 16>), grid3, 256, 0, stream, (const T*)x, ldx, w, bias, act, (T*)y, ldy, n, h, wd)); }
    else { DISPATCH_T(dt, B2U_LAUNCH((conv3x3_smallcin_fwd_kernel<T, 3, 32>), grid3, 256, 0, stream, (const T*)x, ldx, w, bias, act, (T*)y, ldy, n, h, wd)); }
    return B2U_OK;
  }
  if (J > 32) {
    dim3 grid(tiles, n, b2u_cdiv(J, 64));
    DISPATCH_T(dt, B2U_LAUNCH((conv3x3_direct_kernel<T, 64>), grid, 256, 0, stream, (const T*)x, ldx, K, w, dgrad, bias,
                              act, (T*)y, ldy, J, stats, (const T*)mask, ldmask, mask_act, accumulate, n, h, wd));
  } else {
    dim3 grid(tiles, n, b2u_cdiv(J, 32));
    DISPATCH_T(dt, B2U_LAUNCH((conv3x3_direct_kernel<T, 32>), grid, 256, 0, stream, (const T*)x, ldx, K, w, dgrad, bias,
                              act, (T*)y, ldy, J, stats, (const T*)mask, ldmask, mask_act, accumulate, n, h, wd));
  }
  return B2U_OK;
}

int b2u_direct_conv3x3_wgrad(int dt, const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw,
                             float* db, int n, int h, int wd, void* stream) {
  B2U_REQUIRE(n > 0 && h > 0 && wd > 0 && cin > 0 && cout > 0, "conv3x3_wgrad: empty shape");
  if (cin == 1 && lddy % 8 == 0 && (cout == 32 || cout == 16)) {
    long long gl = (long long)n * b2u_cdiv(h, C1_TH) * b2u_cdiv(wd, C1_TW);
    int grid1 = (int)(gl < 2 * B2U_NUM_SMS ? gl : 2 * B2U_NUM_SMS);
    if (cout == 32) { DISPATCH_T(dt, B2U_LAUNCH((conv3x3_c1_wgrad_kernel<T, 32>), grid1, 256, 0, stream, (const T*)x, ldx, (const T*)dy, lddy, dw, db, n, h, wd)); }
    else { DISPATCH_T(dt, B2U_LAUNCH((conv3x3_c1_wgrad_kernel<T, 16>), grid1, 256, 0, stream, (const T*)x, ldx, (const T*)dy, lddy, dw, db, n, h, wd)); }
    return B2U_OK;
  }
  int cib = b2u_cdiv(cin, WG_CI), cob = b2u_cdiv(cout, WG_CO);
  long long ntiles = (long long)n * b2u_cdiv(h, WG_TH) * b2u_cdiv(wd, WG_TW);
  long long nsplit = (4LL * B2U_NUM_SMS * 4) / ((long long)cib * cob);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > ntiles) nsplit = ntiles;
  dim3 grid((unsigned)nsplit, cib, cob);
  DISPATCH_T(dt, B2U_LAUNCH(conv3x3_wgrad_direct_kernel<T>, grid, 128, 0, stream, (const T*)x, ldx, cin, (const T*)dy,
                            lddy, cout, dw, db, n, h, wd, (int)nsplit));
  return B2U_OK;
}

// db[c] += sum over pixels of dy[.][c]  (bias gradient beside the tcgen05 weight-gradient kernel)
// db[c] += sum over pixels of dy[p][c] for either storage type (fallback for producers that cannot emit the sums)
int b2u_channel_sum(int dt, const void* dy, int lddy, int c, long long npix, float* db, void* stream) {
  B2U_REQUIRE(c % 8 == 0 && lddy % 8 == 0 && c <= 2048, "channel_sum: c%%8==0 required");
  int lanes = 256 / (c / 8);
  long long g = (npix + lanes - 1) / lanes;
  if (g > 4LL * B2U_NUM_SMS) g = 4LL * B2U_NUM_SMS;
  DISPATCH_T(dt, B2U_LAUNCH(channel_sum_kernel<T>, (int)g, 256, c * sizeof(float), stream, (const T*)dy, lddy, c, npix, db));
  return B2U_OK;
}

int b2u_channel_sum_f16(const void* dy, int lddy, int c, long long npix, float* db, void* stream) {
  B2U_REQUIRE(c % 8 == 0 && lddy % 8 == 0 && c <= 2048, "channel_sum: c%%8==0 required");
  int lanes = 256 / (c / 8);
  long long g = (npix + lanes - 1) / lanes;
  if (g > 4 * B2U_NUM_SMS) g = 4 * B2U_NUM_SMS;
  B2U_LAUNCH(channel_sum_kernel<__half>, (int)g, 256, c * sizeof(float), stream, (const __half*)dy, lddy, c, npix, db);
  return B2U_OK;
}

int b2u_direct_convt_fwd(int dt, const void* x, int ldx, int cin, const float* w, const float* bias, void* y, int ldy,
                         int cout, int n, int h, int wd, void* stream) {
  long long M = (long long)n * h * wd;
  int Nn = 4 * cout;
  dim3 grid((unsigned)b2u_cdiv(M, 64), b2u_cdiv(Nn, 64), 1);
  DISPATCH_T(dt, {
    CtFwdA<T> la{(const T*)x, ldx};
    CtFwdB lb{w, cin};
    CtFwdEp<T> ep{(T*)y, ldy, cout, h, wd, bias};
    B2U_LAUNCH((gemm64_kernel<CtFwdA<T>, CtFwdB, CtFwdEp<T>>), grid, 256, 0, stream, la, lb, ep, M, Nn, (long long)cin, 1);
  });
  return B2U_OK;
}

int b2u_direct_convt_dgrad(int dt, const void* dy, int lddy, int cout, const float* w, void* dx, int lddx, int cin,
                           const void* mask, int ldmask, int mask_act, int accumulate, int n, int h, int wd,
                           void* stream) {
  long long M = (long long)n * h * wd;
  dim3 grid((unsigned)b2u_cdiv(M, 64), b2u_cdiv(cin, 64), 1);
  DISPATCH_T(dt, {
    CtDgradA<T> la{(const T*)dy, lddy, cout, h, wd};
    CtDgradB lb{w, cin};
    CtDgradEp<T> ep{(T*)dx, lddx, (const T*)mask, ldmask, mask_act, accumulate};
    B2U_LAUNCH((gemm64_kernel<CtDgradA<T>, CtDgradB, CtDgradEp<T>>), grid, 256, 0, stream, la, lb, ep, M, cin,
               (long long)4 * cout, 1);
  });
  return B2U_OK;
}

int b2u_direct_convt_wgrad(int dt, const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw,
                           float* db, int n, int h, int wd, void* stream) {
  long long P = (long long)n * h * wd;
  int Mq = 4 * cout;
  int mb = b2u_cdiv(Mq, 64), nb = b2u_cdiv(cin, 64);
  long long ksplit = (2LL * B2U_NUM_SMS) / ((long long)mb * nb);
  if (ksplit < 1) ksplit = 1;
  if (ksplit > (P + 255) / 256) ksplit = (P + 255) / 256;
  if (ksplit < 1) ksplit = 1;
  dim3 grid(mb, nb, (unsigned)ksplit);
  DISPATCH_T(dt, {
    CtWgradA<T> la{(const T*)dy, lddy, cout, h, wd};
    CtWgradB<T> lb{(const T*)x, ldx};
    CtWgradEp ep{dw, cin};
    B2U_LAUNCH((gemm64_kernel<CtWgradA<T>, CtWgradB<T>, CtWgradEp>), grid, 256, 0, stream, la, lb, ep, (long long)Mq,
               cin, P, (int)ksplit);
    if (db != nullptr) {
      B2U_REQUIRE(cout % 8 == 0 && lddy % 8 == 0, "convt_wgrad: cout%%8==0 required");
      int lanes = 256 / (cout / 8);
      long long opix = 4 * P;
      int g = (int)((opix + lanes - 1) / lanes);
      if (g > 4 * B2U_NUM_SMS) g = 4 * B2U_NUM_SMS;
      B2U_LAUNCH(channel_sum_kernel<T>, g, 256, cout * sizeof(float), stream, (const T*)dy, lddy, cout, opix, db);
    }
  });
  return B2U_OK;
}
