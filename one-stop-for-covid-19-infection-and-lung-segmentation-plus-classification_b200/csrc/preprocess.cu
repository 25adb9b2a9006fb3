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

// ---- P2 + P3: OpenCV's uint8 resize arithmetic, bit for bit ---------------------------------------------
// (modules/imgproc/src/resize.cpp, CV_8UC1; restated in numpy by oracle/cv_resize.py, which tests/ pin against cv2)
//
// Stage 1: for each image and each of its two boxes (x,y,w,h): cv2.resize(crop, (half_w, out_h), INTER_AREA) written side
// by side into mid (out_h x 2*half_w) -- the reference's `cropper` (T1H:236-241) / box re-application (T1H:347-368).
//   both scales >= 1, both integer : resizeAreaFast_   2x2: (a+b+c+d+2) >> 2, else cvRound(sum * (1.f / area))
//   both scales >= 1               : resizeArea_        DecimateAlpha tables, float accumulation IN TABLE ORDER with
//                                                       separate multiply / add roundings (no FMA), cvRound at the end
//   otherwise (a crop smaller than the target in a dimension): the fixed-point bilinear path below with the "area"
//   coefficients  sx = floor(dx*scale), fx = (dx+1) - (sx+1)*inv_scale  (fx <= 0 ? 0 : fx - floor(fx))
// scale = 1. / ((double)dst / src) exactly as cv::resize derives it.
struct AreaTab {            // the DecimateAlpha entries of ONE destination index (computeResizeAreaTab)
  int s_first;              // source index of the first entry
  int n;                    // number of entries (consecutive source indices)
  float a_first, a_mid, a_last;
  bool has_first, has_last;
};

__device__ __forceinline__ AreaTab area_tab(int d, int ssize, double scale) {
  AreaTab t;
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, (double)ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  t.has_first = (double)sx1 - fsx1 > 1e-3;
  t.has_last = fsx2 - (double)sx2 > 1e-3;
  t.a_first = (float)(((double)sx1 - fsx1) / cell);
  t.a_mid = (float)(1.0 / cell);
  t.a_last = (float)(fmin(fmin(fsx2 - (double)sx2, 1.0), cell) / cell);
  t.s_first = t.has_first ? sx1 - 1 : sx1;
  t.n = (sx2 - sx1) + (t.has_first ? 1 : 0) + (t.has_last ? 1 : 0);
  return t;
}

__device__ __forceinline__ float area_alpha(const AreaTab& t, int k) {
  if (k == 0 && t.has_first) return t.a_first;
  if (k == t.n - 1 && t.has_last) return t.a_last;
  return t.a_mid;
}

// 11-bit bilinear coefficients of cv::resize's generic path (ksize 2): source index and the two short weights
__device__ __forceinline__ void linear_coeff(int d, double scale, double inv_scale, bool area_mode, int* s, int* a0, int* a1) {
  float f;
  int si;
  if (!area_mode) {
    f = (float)(((double)d + 0.5) * scale - 0.5);
    si = (int)floorf(f);
    f -= (float)si;
  } else {
    si = (int)floor((double)d * scale);
    f = (float)((double)(d + 1) - (double)(si + 1) * inv_scale);
    f = f <= 0.f ? 0.f : f - floorf(f);
  }
  *s = si;
  *a0 = __float2int_rn(__fmul_rn(1.f - f, 2048.f));       // saturate_cast<short>: cvRound, half to even
  *a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

// horizontal pass of one source row at destination column coefficients (sx, a0, a1): HResizeLinear incl. its borders
__device__ __forceinline__ int linear_row(const uint8_t* row, int sw, int sx, int a0, int a1) {
  if (sx < 0) return (int)row[0] * 2048;
  if (sx >= sw - 1) return (int)row[sw - 1] * 2048;
  return (int)row[sx] * a0 + (int)row[sx + 1] * a1;
}

__device__ __forceinline__ uint8_t linear_fixed_pixel(const uint8_t* src, long long pitch, int sw, int sh, int dx, int dy,
                                                      double scale_x, double inv_x, double scale_y, double inv_y,
                                                      bool area_mode) {
  int sx, ax0, ax1, sy, ay0, ay1;
  linear_coeff(dx, scale_x, inv_x, area_mode, &sx, &ax0, &ax1);
  linear_coeff(dy, scale_y, inv_y, area_mode, &sy, &ay0, &ay1);
  const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);      // rows are clamped, the weights are not
  const int r0 = linear_row(src + (long long)y0 * pitch, sw, sx, ax0, ax1);
  const int r1 = linear_row(src + (long long)y1 * pitch, sw, sx, ax0, ax1);
  const int v = (((ay0 * (r0 >> 4)) >> 16) + ((ay1 * (r1 >> 4)) >> 16) + 2) >> 2;   // VResizeLinear, 8-bit
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__device__ __forceinline__ uint8_t area_pixel(const uint8_t* src, long long pitch, int sw, int sh, int dw, int dh, int dx,
                                              int dy) {
  const double inv_x = (double)dw / (double)sw, inv_y = (double)dh / (double)sh;
  const double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
  if (!(scale_x >= 1.0 && scale_y >= 1.0))
    return linear_fixed_pixel(src, pitch, sw, sh, dx, dy, scale_x, inv_x, scale_y, inv_y, true);
  const int ix = __double2int_rn(scale_x), iy = __double2int_rn(scale_y);
  if (fabs(scale_x - ix) < 2.220446049250313e-16 && fabs(scale_y - iy) < 2.220446049250313e-16) {
    int sum = 0;
    for (int yy = 0; yy < iy; ++yy)
      for (int xx = 0; xx < ix; ++xx) sum += src[(long long)(dy * iy + yy) * pitch + dx * ix + xx];
    if (ix == 2 && iy == 2) return (uint8_t)((sum + 2) >> 2);
    const float sc = __fdiv_rn(1.f, (float)(ix * iy));
    const int q = __float2int_rn(__fmul_rn((float)sum, sc));
    return (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
  const AreaTab tx = area_tab(dx, sw, scale_x), ty = area_tab(dy, sh, scale_y);
  float sum = 0.f;
  for (int j = 0; j < ty.n; ++j) {
    const uint8_t* row = src + (long long)(ty.s_first + j) * pitch;
    float buf = 0.f;
    for (int k = 0; k < tx.n; ++k) buf = __fadd_rn(buf, __fmul_rn((float)row[tx.s_first + k], area_alpha(tx, k)));
    const float term = __fmul_rn(area_alpha(ty, j), buf);
    sum = j == 0 ? term : __fadd_rn(sum, term);
  }
  const int q = __float2int_rn(sum);
  return (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
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
    uint8_t v = 0;
    if (bw > 0 && bh > 0)
      v = area_pixel(in + (long long)n * H * W + (long long)y0 * W + x0, W, bw, bh, half_w, out_h, dx, oy);
    mid[idx] = v;
  }
}

// whole images: cv2.resize(img, (dw, dh), INTER_AREA) -- the NIfTI ingest's 512 x 512 stage (T1H:335)
__global__ void __launch_bounds__(256) area_resize_kernel(const uint8_t* __restrict__ in, int sh, int sw,
                                                          uint8_t* __restrict__ out, int dh, int dw, int N) {
  B2U_PDL_PROLOGUE();
  const long long total = (long long)N * dh * dw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(idx % dw);
    const long long t = idx / dw;
    const int dy = (int)(t % dh);
    const int n = (int)(t / dh);
    out[idx] = area_pixel(in + (long long)n * sh * sw, sw, sw, sh, dw, dh, dx, dy);
  }
}

// ---- CV_64F INTER_AREA + per-slice min-max (the first stage of the NIfTI ingest, T1H:288-297 / 335-337) -----------
// the same three regimes with double accumulators, float32 coefficients promoted to double, separate multiply / add
// roundings, the fast path summing in groups of four like OpenCV's unrolled loop; no rounding at the end
__device__ __forceinline__ double area_pixel_f64(const double* src, long long pitch, int sw, int sh, int dw, int dh, int dx,
                                                 int dy) {
  const double inv_x = (double)dw / (double)sw, inv_y = (double)dh / (double)sh;
  const double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
  if (scale_x >= 1.0 && scale_y >= 1.0) {
    const int ix = __double2int_rn(scale_x), iy = __double2int_rn(scale_y);
    if (fabs(scale_x - ix) < 2.220446049250313e-16 && fabs(scale_y - iy) < 2.220446049250313e-16) {
      const int area = ix * iy;
      const double sc = (double)__fdiv_rn(1.f, (float)area);
      double sum = 0.0, grp = 0.0;
      int k = 0;
      const int full = area & ~3;
      for (int yy = 0; yy < iy; ++yy)
        for (int xx = 0; xx < ix; ++xx, ++k) {
          const double v = src[(long long)(dy * iy + yy) * pitch + dx * ix + xx];
          if (k < full) {
            grp = (k & 3) == 0 ? v : __dadd_rn(grp, v);
            if ((k & 3) == 3) sum = __dadd_rn(sum, grp);
          } else {
            sum = __dadd_rn(sum, v);
          }
        }
      return __dmul_rn(sum, sc);
    }
    const AreaTab tx = area_tab(dx, sw, scale_x), ty = area_tab(dy, sh, scale_y);
    double sum = 0.0;
    for (int j = 0; j < ty.n; ++j) {
      const double* row = src + (long long)(ty.s_first + j) * pitch;
      double buf = 0.0;
      for (int k = 0; k < tx.n; ++k) buf = __dadd_rn(buf, __dmul_rn(row[tx.s_first + k], (double)area_alpha(tx, k)));
      const double term = __dmul_rn((double)area_alpha(ty, j), buf);
      sum = j == 0 ? term : __dadd_rn(sum, term);
    }
    return sum;
  }
  // up-sampling in a dimension: bilinear, "area" coefficients in float32, arithmetic in double
  int sx = (int)floor((double)dx * scale_x);
  float fx = (float)((double)(dx + 1) - (double)(sx + 1) * inv_x);
  fx = fx <= 0.f ? 0.f : fx - floorf(fx);
  if (sx < 0) { fx = 0.f; sx = 0; }
  if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
  int sy = (int)floor((double)dy * scale_y);
  float fy = (float)((double)(dy + 1) - (double)(sy + 1) * inv_y);
  fy = fy <= 0.f ? 0.f : fy - floorf(fy);
  if (sy < 0) { fy = 0.f; sy = 0; }
  if (sy >= sh - 1) { fy = 0.f; sy = sh - 1; }
  const int sx1 = min(sx + 1, sw - 1), sy1 = min(sy + 1, sh - 1);
  const double a0 = (double)(1.f - fx), a1 = (double)fx, b0 = (double)(1.f - fy), b1 = (double)fy;
  const double* r0 = src + (long long)sy * pitch;
  const double* r1 = src + (long long)sy1 * pitch;
  const bool edge = sx >= sw - 1;                                   // (dx >= xmax: D = S[sx] * ONE)
  const double h0 = edge ? r0[sx] : __dadd_rn(__dmul_rn(r0[sx], a0), __dmul_rn(r0[sx1], a1));
  const double h1 = edge ? r1[sx] : __dadd_rn(__dmul_rn(r1[sx], a0), __dmul_rn(r1[sx1], a1));
  return __dadd_rn(__dmul_rn(h0, b0), __dmul_rn(h1, b1));
}

__global__ void __launch_bounds__(256) area_resize_f64_kernel(const double* __restrict__ in, int sh, int sw,
                                                              double* __restrict__ out, int dh, int dw, int N) {
  B2U_PDL_PROLOGUE();
  const long long total = (long long)N * dh * dw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(idx % dw);
    const long long t = idx / dw;
    const int dy = (int)(t % dh);
    const int n = (int)(t / dh);
    const double* src = in + (long long)n * sh * sw;
    out[idx] = (sh == dh && sw == dw) ? src[(long long)dy * sw + dx]          // cv::resize: same size = copy
                                      : area_pixel_f64(src, sw, sw, sh, dw, dh, dx, dy);
  }
}

// (img - min) / (max - min) per image in double, exactly numpy's two roundings; a constant image gives 0/0 = NaN like
// the reference (T1H:337).  One block per image.
__global__ void __launch_bounds__(1024) minmax_normalize_f64_kernel(double* __restrict__ img, long long count) {
  B2U_PDL_PROLOGUE();
  double* p = img + (long long)blockIdx.x * count;
  __shared__ double smin[32], smax[32];
  double lo = INFINITY, hi = -INFINITY;
  bool nan = false;
  for (long long i = threadIdx.x; i < count; i += blockDim.x) {
    const double v = p[i];
    nan |= v != v;
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  for (int o = 16; o >= 1; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  nan = __syncthreads_or(nan);
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  lo = smin[0];
  hi = smax[0];
  for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { lo = fmin(lo, smin[k]); hi = fmax(hi, smax[k]); }
  if (nan) { lo = NAN; hi = NAN; }                                 // numpy's min / max propagate NaN
  const double den = __dsub_rn(hi, lo);
  for (long long i = threadIdx.x; i < count; i += blockDim.x) p[i] = __ddiv_rn(__dsub_rn(p[i], lo), den);
}

// Stage 2: cv2.resize(mid, (final, final), INTER_LINEAR) for uint8 (OpenCV fixed point: 11-bit coefficients, two-pass
// with the (>>4, >>16, +2 >>2) rounding of its 8-bit VResizeLinear), then np.uint8(.) / 255 (T1H:485-488).
__global__ void __launch_bounds__(256) linear_resize_scale_kernel(const uint8_t* __restrict__ mid, int mh, int mw, int fd,
                                                                  float* __restrict__ out, uint8_t* __restrict__ out_u8,
                                                                  int N) {
  B2U_PDL_PROLOGUE();
  const long long total = (long long)N * fd * fd;
  const double inv_x = (double)fd / (double)mw, inv_y = (double)fd / (double)mh;
  const double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(idx % fd);
    const long long t = idx / fd;
    const int dy = (int)(t % fd);
    const int n = (int)(t / fd);
    const uint8_t q = linear_fixed_pixel(mid + (long long)n * mh * mw, mw, mw, mh, dx, dy, scale_x, inv_x, scale_y, inv_y, false);
    if (out != nullptr) out[idx] = __fdiv_rn((float)q, 255.0f);
    if (out_u8 != nullptr) out_u8[idx] = q;
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
  B2U_LAUNCH(linear_resize_scale_kernel, g2, 256, 0, stream, (const uint8_t*)mid_u8, out_h, 2 * half_w, final_dim, out,
             (uint8_t*)nullptr, n);
  return B2U_OK;
}

// cv2.resize(img, (dst_w, dst_h), interpolation) for n uint8 images of src_h x src_w; interpolation 1 = INTER_LINEAR,
// 3 = INTER_AREA (OpenCV's enum values), bit-exact against OpenCV (tests/test_gpu_preprocess.py)
extern "C" int b2u_resize_u8(const uint8_t* in, int n, int src_h, int src_w, uint8_t* out, int dst_h, int dst_w,
                             int interpolation, void* stream) {
  B2U_REQUIRE(in != nullptr && out != nullptr && n > 0 && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0, "resize_u8: args");
  long long total = (long long)n * dst_h * dst_w;
  int grid = (int)((total + 255) / 256);
  if (grid > 16 * B2U_NUM_SMS) grid = 16 * B2U_NUM_SMS;
  if (interpolation == 3) {
    B2U_LAUNCH(area_resize_kernel, grid, 256, 0, stream, in, src_h, src_w, out, dst_h, dst_w, n);
  } else if (interpolation == 1) {
    B2U_REQUIRE(dst_h == dst_w, "resize_u8: INTER_LINEAR is implemented for square targets (the reference's new_dim)");
    B2U_LAUNCH(linear_resize_scale_kernel, grid, 256, 0, stream, in, src_h, src_w, dst_h, (float*)nullptr, out, n);
  } else {
    b2u_set_error("resize_u8: interpolation %d not supported (1 = INTER_LINEAR, 3 = INTER_AREA)", interpolation);
    return B2U_ERR_ARG;
  }
  return B2U_OK;
}

// the first stage of the NIfTI ingest on the device: cv2.resize(slice, (dst_w, dst_h), INTER_AREA) on float64 slices
// (T1H:335), bit-exact against OpenCV's CV_64F arithmetic; then optionally (img - min) / (max - min) per slice (T1H:336-337)
extern "C" int b2u_resize_area_f64(const double* in, int n, int src_h, int src_w, double* out, int dst_h, int dst_w,
                                   int minmax_normalize, void* stream) {
  B2U_REQUIRE(in != nullptr && out != nullptr && n > 0 && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0, "resize_area_f64: args");
  long long total = (long long)n * dst_h * dst_w;
  int grid = (int)((total + 255) / 256);
  if (grid > 32 * B2U_NUM_SMS) grid = 32 * B2U_NUM_SMS;
  B2U_LAUNCH(area_resize_f64_kernel, grid, 256, 0, stream, in, src_h, src_w, out, dst_h, dst_w, n);
  if (minmax_normalize) B2U_LAUNCH(minmax_normalize_f64_kernel, n, 1024, 0, stream, out, (long long)dst_h * dst_w);
  return B2U_OK;
}
