// Preprocessing kernels: CLAHE, lung-box crop + area resize + hconcat + bilinear
// resize + /255.  uint8 / integer work, HBM-bound; coalesced row-wise access.
//
// P1 restates OpenCV's CLAHE (cv2.createCLAHE(clipLimit, (tiles,tiles)).apply, called at
// /root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:169-170) operation for operation
// in fp32 WITHOUT fused multiply-adds, so the result is bit-identical to cv2 (tests/test_gpu_preprocess.py):
//   per tile: 256-bin histogram -> clip at max(1, int(clip*area/256)) -> uniform redistribution + residual
//   stepping -> cumulative sum -> LUT = cvRound(sum * (255/area));
//   per pixel: bilinear blend of the four surrounding tile LUTs.
#include "common.cuh"
#include "launch.cuh"

namespace {

__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ lut, int H,
                                                        int W, int tiles, int clip_limit, float lut_scale) {
  B2U_PDL_PROLOGUE();
  // one block per (image, tile_y, tile_x); thread i owns histogram bin i
  __shared__ int hist[256];
  __shared__ int scan[256];
  __shared__ int s_clipped;
  const int tx = blockIdx.x % tiles, ty = (blockIdx.x / tiles) % tiles, n = blockIdx.x / (tiles * tiles);
  const int th = H / tiles, tw = W / tiles;
  const int i = threadIdx.x;
  hist[i] = 0;
  if (i == 0) s_clipped = 0;
  __syncthreads();
  const uint8_t* base = in + ((long long)n * H + (long long)ty * th) * W + (long long)tx * tw;
  for (int p = i; p < th * tw; p += 256) {
    int r = p / tw, c = p - r * tw;
    atomicAdd(&hist[base[(long long)r * W + c]], 1);
  }
  __syncthreads();
  int hv = hist[i];
  if (clip_limit > 0) {
    int over = hv > clip_limit ? hv - clip_limit : 0;
    if (over > 0) { atomicAdd(&s_clipped, over); hv = clip_limit; }
    __syncthreads();
    const int clipped = s_clipped;
    const int batch = clipped / 256;
    int resid = clipped - batch * 256;
    hv += batch;
    if (resid != 0) {
      int step = 256 / resid;
      if (step < 1) step = 1;
      if (i % step == 0 && i / step < resid) hv += 1;
    }
  }
  // inclusive scan over 256 bins (Hillis-Steele in shared memory)
  scan[i] = hv;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    int v = scan[i];
    if (i >= off) v += scan[i - off];
    __syncthreads();
    scan[i] = v;
    __syncthreads();
  }
  float f = __fmul_rn((float)scan[i], lut_scale);
  int q = __float2int_rn(f);
  q = q < 0 ? 0 : (q > 255 ? 255 : q);
  lut[(long long)blockIdx.x * 256 + i] = (uint8_t)q;
}

__global__ void __launch_bounds__(256) clahe_interp_kernel(const uint8_t* __restrict__ in, const uint8_t* __restrict__ lut,
                                                           uint8_t* __restrict__ out, int N, int H, int W, int tiles) {
  B2U_PDL_PROLOGUE();
  const int th = H / tiles, tw = W / tiles;
  const float inv_tw = __fdiv_rn(1.0f, (float)tw), inv_th = __fdiv_rn(1.0f, (float)th);
  const long long total = (long long)N * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const long long t = idx / W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
    int tx1 = (int)floorf(txf);
    const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
    int tx2 = tx1 + 1;
    tx1 = tx1 < 0 ? 0 : tx1;
    tx2 = tx2 > tiles - 1 ? tiles - 1 : tx2;
    const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    int ty1 = (int)floorf(tyf);
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    int ty2 = ty1 + 1;
    ty1 = ty1 < 0 ? 0 : ty1;
    ty2 = ty2 > tiles - 1 ? tiles - 1 : ty2;
    const int v = in[idx];
    const uint8_t* l = lut + (long long)n * tiles * tiles * 256;
    const float l11 = (float)l[(ty1 * tiles + tx1) * 256 + v], l12 = (float)l[(ty1 * tiles + tx2) * 256 + v];
    const float l21 = (float)l[(ty2 * tiles + tx1) * 256 + v], l22 = (float)l[(ty2 * tiles + tx2) * 256 + v];
    const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
    const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    int q = __float2int_rn(res);
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    out[idx] = (uint8_t)q;
  }
}

// ---- P2 + P3 ---------------------------------------------------------------------------------------
// Stage 1: for each image and each of its two boxes (x,y,w,h): cv2.resize(crop, (half_w, out_h), INTER_AREA)
// written side by side into mid (out_h x 2*half_w) -- the reference's `cropper` (T1H:236-241) / box
// re-application (T1H:347-368).  INTER_AREA = exact box filter with fractional pixel coverage (float
// accumulation like OpenCV's resizeArea_; integer scale factors reduce to plain averaging), rounded to uint8.
// When a crop is SMALLER than the target in a dimension OpenCV switches that call to its bilinear path; the
// same rule is applied here per call (area coefficients: fx = (dx+1) - (sx+1)/scale, clamped).
__device__ __forceinline__ float area_axis_weight(int s, float lo, float hi) {
  float a = fmaxf((float)s, lo), b = fminf((float)(s + 1), hi);
  return fmaxf(b - a, 0.f);
}

__global__ void __launch_bounds__(256) crop_area_resize_kernel(const uint8_t* __restrict__ in, int H, int W,
                                                               const int* __restrict__ boxes, int half_w, int out_h,
                                                               uint8_t* __restrict__ mid, int N) {
  B2U_PDL_PROLOGUE();
  const int ow = 2 * half_w;
  const long long total = (long long)N * out_h * ow;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % ow);
    const long long t = idx / ow;
    const int oy = (int)(t % out_h);
    const int n = (int)(t / out_h);
    const int side = ox >= half_w;
    const int dx = ox - side * half_w;
    const int* bx = boxes + (long long)n * 8 + side * 4;
    const int x0 = bx[0], y0 = bx[1], bw = bx[2], bh = bx[3];
    const uint8_t* src = in + (long long)n * H * W;
    float val;
    if (bw <= 0 || bh <= 0) {
      val = 0.f;
    } else {
      const float sx = (float)bw / (float)half_w, sy = (float)bh / (float)out_h;
      if (sx >= 1.f && sy >= 1.f) {
        // area (box) filter with fractional coverage
        const float fx0 = dx * sx, fx1 = fx0 + sx, fy0 = oy * sy, fy1 = fy0 + sy;
        int ix0 = (int)floorf(fx0), ix1 = min((int)ceilf(fx1), bw), iy0 = (int)floorf(fy0), iy1 = min((int)ceilf(fy1), bh);
        float acc = 0.f, wsum = 0.f;
        for (int yy = iy0; yy < iy1; ++yy) {
          const float wy = area_axis_weight(yy, fy0, fy1);
          if (wy <= 0.f) continue;
          float row = 0.f, wrow = 0.f;
          for (int xx = ix0; xx < ix1; ++xx) {
            const float wx = area_axis_weight(xx, fx0, fx1);
            row += wx * (float)src[(long long)(y0 + yy) * W + x0 + xx];
            wrow += wx;
          }
          acc += wy * row;
          wsum += wy * wrow;
        }
        val = wsum > 0.f ? acc / wsum : 0.f;
      } else {
        // OpenCV's INTER_AREA up-scaling rule: bilinear with area-style coefficients
        int sx0 = (int)floorf(dx * sx);
        float fx = (float)(dx + 1) - (float)(sx0 + 1) / sx;
        fx = fx <= 0.f ? 0.f : fx - floorf(fx);
        int sy0 = (int)floorf(oy * sy);
        float fy = (float)(oy + 1) - (float)(sy0 + 1) / sy;
        fy = fy <= 0.f ? 0.f : fy - floorf(fy);
        int sx1 = min(sx0 + 1, bw - 1), sy1 = min(sy0 + 1, bh - 1);
        sx0 = min(sx0, bw - 1);
        sy0 = min(sy0, bh - 1);
        const float p00 = src[(long long)(y0 + sy0) * W + x0 + sx0], p01 = src[(long long)(y0 + sy0) * W + x0 + sx1];
        const float p10 = src[(long long)(y0 + sy1) * W + x0 + sx0], p11 = src[(long long)(y0 + sy1) * W + x0 + sx1];
        val = (p00 * (1.f - fx) + p01 * fx) * (1.f - fy) + (p10 * (1.f - fx) + p11 * fx) * fy;
      }
    }
    int q = __float2int_rn(val);
    mid[idx] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
}

// Stage 2: cv2.resize(mid, (final, final), INTER_LINEAR) for uint8 (OpenCV fixed point: 11-bit
// coefficients, two-pass with the (>>4, >>16, +2 >>2) rounding of its 8-bit VResizeLinear) then /255.
__global__ void __launch_bounds__(256) linear_resize_scale_kernel(const uint8_t* __restrict__ mid, int mh, int mw, int fd,
                                                                  float* __restrict__ out, int N) {
  B2U_PDL_PROLOGUE();
  const long long total = (long long)N * fd * fd;
  const float scx = (float)mw / (float)fd, scy = (float)mh / (float)fd;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(idx % fd);
    const long long t = idx / fd;
    const int dy = (int)(t % fd);
    const int n = (int)(t / fd);
    float fx = (float)((dx + 0.5) * (double)scx - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= mw - 1) { fx = 0.f; sx = mw - 1; }
    float fy = (float)((dy + 0.5) * (double)scy - 0.5);
    int sy = (int)floorf(fy);
    fy -= sy;
    if (sy < 0) { fy = 0.f; sy = 0; }
    if (sy >= mh - 1) { fy = 0.f; sy = mh - 1; }
    const int sx1 = min(sx + 1, mw - 1), sy1 = min(sy + 1, mh - 1);
    // saturate_cast<short>(f * 2048): round to nearest even
    const int ax1 = __float2int_rn(fx * 2048.f), ax0 = __float2int_rn((1.f - fx) * 2048.f);
    const int ay1 = __float2int_rn(fy * 2048.f), ay0 = __float2int_rn((1.f - fy) * 2048.f);
    const uint8_t* src = mid + (long long)n * mh * mw;
    const int r0 = src[(long long)sy * mw + sx] * ax0 + src[(long long)sy * mw + sx1] * ax1;     // horizontal pass
    const int r1 = src[(long long)sy1 * mw + sx] * ax0 + src[(long long)sy1 * mw + sx1] * ax1;
    const int v = ((((ay0 * (r0 >> 4)) >> 16) + ((ay1 * (r1 >> 4)) >> 16) + 2) >> 2);
    const int q = v < 0 ? 0 : (v > 255 ? 255 : v);
    out[idx] = (float)q / 255.0f;
  }
}

}  // namespace

extern "C" int b2u_clahe_u8(const uint8_t* in, uint8_t* out, int n, int h, int wd, float clip_limit, int tiles,
                            void* ws, size_t ws_bytes, void* stream) {
  B2U_REQUIRE(n > 0 && h > 0 && wd > 0 && tiles > 0, "clahe: empty input");
  B2U_REQUIRE(h % tiles == 0 && wd % tiles == 0, "clahe: image %dx%d must be divisible by the %d-tile grid", h, wd, tiles);
  size_t need = (size_t)n * tiles * tiles * 256;
  B2U_REQUIRE(ws != nullptr && ws_bytes >= need, "clahe: workspace needs %zu bytes", need);
  const int area = (h / tiles) * (wd / tiles);
  int clip = 0;
  if (clip_limit > 0.f) {
    clip = (int)(clip_limit * area / 256);
    if (clip < 1) clip = 1;
  }
  const float lut_scale = 255.0f / (float)area;
  B2U_LAUNCH(clahe_lut_kernel, n * tiles * tiles, 256, 0, stream, in, (uint8_t*)ws, h, wd, tiles, clip, lut_scale);
  long long total = (long long)n * h * wd;
  int grid = (int)((total + 255) / 256);
  if (grid > 16 * B2U_NUM_SMS) grid = 16 * B2U_NUM_SMS;
  B2U_LAUNCH(clahe_interp_kernel, grid, 256, 0, stream, in, (const uint8_t*)ws, out, n, h, wd, tiles);
  return B2U_OK;
}

extern "C" int b2u_crop_resize(const uint8_t* in, int n, int h, int wd, const int* boxes, int half_w, int out_h,
                               int final_dim, uint8_t* mid_u8, float* out, void* stream) {
  B2U_REQUIRE(n > 0 && h > 0 && wd > 0 && half_w > 0 && out_h > 0 && final_dim > 0, "crop_resize: empty input");
  B2U_REQUIRE(mid_u8 != nullptr && out != nullptr && boxes != nullptr, "crop_resize: null buffer");
  long long t1 = (long long)n * out_h * 2 * half_w, t2 = (long long)n * final_dim * final_dim;
  int g1 = (int)((t1 + 255) / 256), g2 = (int)((t2 + 255) / 256);
  if (g1 > 16 * B2U_NUM_SMS) g1 = 16 * B2U_NUM_SMS;
  if (g2 > 16 * B2U_NUM_SMS) g2 = 16 * B2U_NUM_SMS;
  B2U_LAUNCH(crop_area_resize_kernel, g1, 256, 0, stream, in, h, wd, boxes, half_w, out_h, mid_u8, n);
  B2U_LAUNCH(linear_resize_scale_kernel, g2, 256, 0, stream, (const uint8_t*)mid_u8, out_h, 2 * half_w, final_dim, out, n);
  return B2U_OK;
}
