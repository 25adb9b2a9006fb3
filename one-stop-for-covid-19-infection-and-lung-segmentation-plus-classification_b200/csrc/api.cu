// C-ABI glue: error string, conv dispatch (tcgen05 tensor path vs exact CUDA-core path), the op-list
// executor with CUDA-graph capture, and the NCCL gradient exchange.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "launch.cuh"
#include "internal.h"

unsigned long long g_b2u_launches = 0;
int g_b2u_pdl = 0;          // programmatic dependent launch (launch.cuh); measured 4 % slower on the U-Net step: off

static thread_local char g_err[1024] = "";

void b2u_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int b2u_version(void) { return B2U_VERSION; }
extern "C" const char* b2u_last_error(void) { return g_err; }
extern "C" size_t b2u_ws_bytes(void) { return (size_t)64 << 20; }
extern "C" long long b2u_launch_count(void) { return (long long)__atomic_load_n(&g_b2u_launches, __ATOMIC_RELAXED); }

static int g_tc_state = -1;   // -1 unknown, 0 off, 1 on
extern "C" int b2u_tensor_path_available(void) {
  if (g_tc_state < 0) {
    int dev = 0;
    cudaDeviceProp prop;
    const char* env = getenv("B2U_DISABLE_TC");
    if (env != nullptr && env[0] == '1') g_tc_state = 0;
    else if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) g_tc_state = 0;
    else g_tc_state = (prop.major == 10 && b2u_tc_compiled()) ? 1 : 0;
  }
  return g_tc_state;
}

extern int g_b2u_side_stream;
extern int g_b2u_comm_overlap;
static int side_init();
extern "C" int b2u_set_option(const char* name, int value) {
  if (name == nullptr) return -1;
  if (strcmp(name, "tc_halo") == 0) {
    int old = g_b2u_tc_halo;
    g_b2u_tc_halo = value;
    return old;
  }
  if (strcmp(name, "l2_fetch") == 0) {
    // DRAM -> L2 fetch granularity of the device (cudaLimitMaxL2FetchGranularity: 32 / 64 / 128 bytes).  Kernels that read
    // one half of a 128-byte concat pixel (max-pool backward, transposed-conv gradients) pay for the other half at 128.
    size_t old = 0;
    cudaDeviceGetLimit(&old, cudaLimitMaxL2FetchGranularity);
    if (value > 0 && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value) != cudaSuccess) {
      cudaGetLastError();
      return -1;
    }
    return (int)old;
  }
  if (strcmp(name, "pdl") == 0) {
    int old = g_b2u_pdl;
    g_b2u_pdl = value ? 1 : 0;
    return old;
  }
  if (strcmp(name, "tc_dwmerge") == 0) {
    int old = g_b2u_tc_dwmerge;
    g_b2u_tc_dwmerge = value;
    return old;
  }
  if (strcmp(name, "tc_rowstrip") == 0) {
    int old = g_b2u_tc_rowstrip;
    g_b2u_tc_rowstrip = value;
    return old;
  }
  if (strcmp(name, "tc_dw_packed") == 0) {
    int old = g_b2u_tc_dw_packed;
    g_b2u_tc_dw_packed = value ? 1 : 0;
    return old;
  }
  if (strcmp(name, "tc_dw_epi8") == 0) {
    int old = g_b2u_tc_dw_epi8;
    g_b2u_tc_dw_epi8 = value ? 1 : 0;
    return old;
  }
  if (strcmp(name, "tc_2sm_max_j") == 0) {
    int old = g_b2u_tc_2sm_max_j;
    g_b2u_tc_2sm_max_j = value;
    return old;
  }
  if (strcmp(name, "tc_3sm") == 0) {
    int old = g_b2u_tc_3sm;
    g_b2u_tc_3sm = value ? 1 : 0;
    return old;
  }
  if (strcmp(name, "side_stream") == 0) {
    int old = g_b2u_side_stream;
    if (value && side_init() != B2U_OK) return -1;        // create the stream / events outside any stream capture
    g_b2u_side_stream = value ? 1 : 0;
    return old;
  }
  if (strcmp(name, "comm_overlap") == 0) {
    int old = g_b2u_comm_overlap;
    if (value && side_init() != B2U_OK) return -1;
    g_b2u_comm_overlap = value ? 1 : 0;
    return old;
  }
  if (strcmp(name, "bn_async") == 0) {
    int old = g_b2u_bn_async;
    g_b2u_bn_async = value;
    return old;
  }
  if (strcmp(name, "tc_mcast") == 0) {
    int old = g_b2u_tc_mcast;
    g_b2u_tc_mcast = value;
    return old;
  }
  if (strcmp(name, "tc_max_ctas") == 0) {
    int old = g_b2u_tc_max_ctas;
    g_b2u_tc_max_ctas = value;
    return old;
  }
  if (strcmp(name, "tc_bgroup") == 0) {
    int old = g_b2u_tc_bgroup;
    g_b2u_tc_bgroup = value;
    return old;
  }
  if (strcmp(name, "convt_jt") == 0) {
    int old = g_b2u_convt_jt;
    g_b2u_convt_jt = value;
    return old;
  }
  if (strcmp(name, "wgrad_dhm") == 0) {
    int old = g_b2u_wgrad_dhm;
    g_b2u_wgrad_dhm = value;
    return old;
  }
  if (strcmp(name, "wgrad_halo") == 0) {
    int old = g_b2u_wgrad_halo;
    g_b2u_wgrad_halo = value;
    return old;
  }
  if (strcmp(name, "tc_debug") == 0) {
    if (value && g_b2u_dbg == nullptr) {
      if (cudaMalloc(&g_b2u_dbg, 64 * 8 * sizeof(long long)) != cudaSuccess) return -1;
      cudaMemset(g_b2u_dbg, 0, 64 * 8 * sizeof(long long));
    } else if (!value && g_b2u_dbg != nullptr) {
      cudaFree(g_b2u_dbg);
      g_b2u_dbg = nullptr;
    }
    return 0;
  }
  if (strcmp(name, "tc_debug_read") == 0) {   // value = host pointer low bits are not passable: see b2u_debug_read
    return g_b2u_dbg != nullptr;
  }
  if (strcmp(name, "tensor_path") == 0) {
    int old = b2u_tensor_path_available();
    g_tc_state = value ? (b2u_tc_compiled() ? 1 : 0) : 0;
    return old;
  }
  return -1;
}

extern "C" int b2u_debug_read(long long* h_out, int count) {
  if (g_b2u_dbg == nullptr || count > 64 * 8) return -1;
  return cudaMemcpy(h_out, g_b2u_dbg, count * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

// ------------------------------------------------------------------------------------------
// conv dispatch
// ------------------------------------------------------------------------------------------
// Thin layers (N = Cout <= 64): the dw-merged kernel (conv_tc3w.cu) reads the A operand a third as often as the halo
// kernel.  tc_dwmerge = 1 takes every shape the kernel supports (A/B runs, tests), 2 only the shapes where it measured
// faster on B200 (tools/ab_ops.py --opt tc_dwmerge=0,1; profiles/).
// tc_dwmerge = 3: additionally every eligible layer of a TRAINING op list (forward ops flagged by the planner, all data
// gradients), whose epilogue may shuffle its shifted partial sums as fp16 pairs ("tc_dw_packed", conv_tc3w.cu).
int g_b2u_tc_dw_packed = 1;
static bool use_dwmerge(int K, int J, int h, int wd, int training) {
  if (g_b2u_tc_dwmerge == 0 || !b2u_tc_conv3x3_dwmerge_ok(K, J)) return false;
  if (g_b2u_tc_dwmerge == 1) return true;
  if (K >= 128 && (long long)h * wd >= 128 * 128) return true;
  return g_b2u_tc_dwmerge == 3 && training && (long long)h * wd >= 128 * 128;
}

// Row-strip kernel (conv_tc3r.cu, Cout = 16 / 32).  Measured on B200 (tools/one_op.py, profiles/NOTES_r2.md): a tcgen05.mma
// of M = 128 costs about 25 + 1.2 N cycles here whatever K slab it reads, so merging three taps into N = 96 saves only the
// fixed part (150 clk against 3 x 62) while the kernel gives up the two-CTA interleaving of the halo kernel: slower at
// K = 32 / 64 (0.093 against 0.081 ms for 32 -> 32 at 512^2), faster once the MMA phase dominates (128 -> 32 at 512^2:
// 0.181 against 0.241 ms; U-Net++'s level-1 nodes with 96 / 128 input channels).
static bool use_rowstrip(int K, int J, int h, int wd) {
  if (g_b2u_tc_rowstrip == 0 || !b2u_tc_conv3x3_rowstrip_ok(K, J, wd)) return false;
  if (g_b2u_tc_rowstrip == 1) return true;
  return K >= 96 && wd >= 128 && (long long)h * wd >= 128 * 128;
}

// `relu_bits` (optional, op lists only): packed 1-bit mask of y > 0, written by the halo kernel's epilogue or, on the
// other paths, by one extra pass over y
static int conv3x3_fwd_wp(int dt, const void* x, int ldx, int cin, const float* w, const float* bias, int act, void* y,
                          int ldy, int cout, double* stats, int n, int h, int wd, void* ws, size_t ws_bytes,
                          const void* wp, void* relu_bits, void* stream, int training = 0, int k_src = 0,
                          const float* post_scale = nullptr, const float* post_shift = nullptr) {
  int rc;
  if (k_src != 0)       // zero-padded input tensor: `w` has only k_src input channels, only the prepacked copy matches
    B2U_REQUIRE(wp != nullptr && dt == B2U_F16 && b2u_tensor_path_available() && b2u_tc_conv3x3_ok(cin, cout, ldx, ldy),
                "conv3x3_fwd: a channel-padded input needs the tensor path and prepacked weights");
  if (post_scale != nullptr || post_shift != nullptr) {
    // inference plans: the BatchNormalization that follows the activation, as a per-channel affine in the epilogue
    B2U_REQUIRE(dt == B2U_F16 && b2u_tensor_path_available() && b2u_tc_conv3x3_ok(cin, cout, ldx, ldy) && stats == nullptr &&
                relu_bits == nullptr && post_scale != nullptr && post_shift != nullptr,
                "conv3x3_fwd: the post-activation affine needs the tensor path of a plain forward");
    return b2u_tc_conv3x3_halo(x, ldx, cin, w, 0, bias, act, y, ldy, cout, nullptr, nullptr, nullptr, 0, 0, 0, n, h, wd, ws,
                               ws_bytes, wp, stream, nullptr, post_scale, post_shift);
  }
  if (dt == B2U_F16 && b2u_tensor_path_available() && b2u_tc_conv3x3_ok(cin, cout, ldx, ldy)) {
    if (use_rowstrip(cin, cout, h, wd))
      return b2u_tc_conv3x3_rowstrip(x, ldx, cin, w, 0, bias, act, y, ldy, cout, stats, nullptr, nullptr, 0, 0, 0, n, h, wd, ws,
                                     ws_bytes, wp, stream, relu_bits);
    if (use_dwmerge(cin, cout, h, wd, training))
      return b2u_tc_conv3x3_dwmerge(x, ldx, cin, w, 0, bias, act, y, ldy, cout, stats, nullptr, nullptr, 0, 0, 0, n, h, wd,
                                    ws, ws_bytes, wp, stream, relu_bits, training && g_b2u_tc_dw_packed);
    if (g_b2u_tc_halo)
      return b2u_tc_conv3x3_halo(x, ldx, cin, w, 0, bias, act, y, ldy, cout, stats, nullptr, nullptr, 0, 0, 0, n, h, wd,
                                 ws, ws_bytes, wp, stream, relu_bits);
    rc = b2u_tc_conv3x3(x, ldx, cin, w, 0, bias, act, y, ldy, cout, stats, nullptr, nullptr, 0, 0, 0, n, h, wd, ws,
                        ws_bytes, wp, stream, nullptr);
  } else {
    int bits_done = 0;
    rc = b2u_direct_conv3x3(dt, x, ldx, cin, w, 0, bias, act, y, ldy, cout, stats, nullptr, 0, 0, 0, n, h, wd, stream,
                            relu_bits, &bits_done);
    if (bits_done) return rc;
  }
  if (rc != B2U_OK || relu_bits == nullptr) return rc;
  return b2u_relu_bits(dt, y, ldy, cout, (long long)n * h * wd, relu_bits, stream);
}

extern "C" int b2u_conv3x3_fwd(int dt, const void* x, int ldx, int cin, const float* w, const float* bias, int act,
                               void* y, int ldy, int cout, double* stats, int n, int h, int wd, void* ws,
                               size_t ws_bytes, void* stream) {
  return conv3x3_fwd_wp(dt, x, ldx, cin, w, bias, act, y, ldy, cout, stats, n, h, wd, ws, ws_bytes, nullptr, nullptr,
                        stream);
}

// data gradient + optional `colsum` (op lists only): colsum[c] += sum over pixels of the dx values written, i.e. the
// bias gradient of the layer that produced the tensor dx belongs to
static int conv3x3_dgrad_cs(int dt, const void* dy, int lddy, int cout, const float* w, void* dx, int lddx, int cin,
                            const void* mask, int ldmask, int mask_act, int accumulate, float* colsum, int n, int h,
                            int wd, void* ws, size_t ws_bytes, const void* wp, void* stream) {
  const bool tc = dt == B2U_F16 && b2u_tensor_path_available() && b2u_tc_conv3x3_ok(cout, cin, lddy, lddx);
  const bool bits = mask != nullptr && mask_act == B2U_ACT_RELU_BITS;
  if (tc && !(bits && accumulate) && use_rowstrip(cout, cin, h, wd))
    return b2u_tc_conv3x3_rowstrip(dy, lddy, cout, w, 1, nullptr, B2U_ACT_NONE, dx, lddx, cin, nullptr, colsum, mask, ldmask,
                                   mask_act, accumulate, n, h, wd, ws, ws_bytes, wp, stream, nullptr);
  if (tc && !(bits && accumulate) && use_dwmerge(cout, cin, h, wd, 1))
    return b2u_tc_conv3x3_dwmerge(dy, lddy, cout, w, 1, nullptr, B2U_ACT_NONE, dx, lddx, cin, nullptr, colsum, mask, ldmask,
                                  mask_act, accumulate, n, h, wd, ws, ws_bytes, wp, stream, nullptr, g_b2u_tc_dw_packed);
  if (tc && g_b2u_tc_halo)
    return b2u_tc_conv3x3_halo(dy, lddy, cout, w, 1, nullptr, B2U_ACT_NONE, dx, lddx, cin, nullptr, colsum, mask, ldmask, mask_act,
                               accumulate, n, h, wd, ws, ws_bytes, wp, stream, nullptr);
  if (tc && !bits)
    return b2u_tc_conv3x3(dy, lddy, cout, w, 1, nullptr, B2U_ACT_NONE, dx, lddx, cin, nullptr, colsum, mask, ldmask, mask_act,
                          accumulate, n, h, wd, ws, ws_bytes, wp, stream, nullptr);
  int rc;
  if (bits) {
    // 1-bit mask on a path whose kernel cannot read it: unmasked data gradient, then one masking pass
    B2U_REQUIRE(!accumulate, "conv3x3_dgrad: a 1-bit mask cannot be combined with accumulate");
    rc = tc ? b2u_tc_conv3x3(dy, lddy, cout, w, 1, nullptr, B2U_ACT_NONE, dx, lddx, cin, nullptr, nullptr, nullptr, 0, 0, 0, n,
                             h, wd, ws, ws_bytes, wp, stream, nullptr)
            : b2u_direct_conv3x3(dt, dy, lddy, cout, w, 1, nullptr, B2U_ACT_NONE, dx, lddx, cin, nullptr, nullptr, 0, 0, 0, n,
                                 h, wd, stream);
    if (rc == B2U_OK) rc = b2u_apply_relu_bits(dt, dx, lddx, cin, (long long)n * h * wd, mask, stream);
  } else {
    rc = b2u_direct_conv3x3(dt, dy, lddy, cout, w, 1, nullptr, B2U_ACT_NONE, dx, lddx, cin, nullptr, mask, ldmask,
                            mask_act, accumulate, n, h, wd, stream);
  }
  if (rc != B2U_OK || colsum == nullptr) return rc;
  return b2u_channel_sum(dt, dx, lddx, cin, (long long)n * h * wd, colsum, stream);       // exact path: extra pass
}

extern "C" int b2u_conv3x3_dgrad(int dt, const void* dy, int lddy, int cout, const float* w, void* dx, int lddx,
                                 int cin, const void* mask, int ldmask, int mask_act, int accumulate, int n, int h,
                                 int wd, void* ws, size_t ws_bytes, void* stream) {
  return conv3x3_dgrad_cs(dt, dy, lddy, cout, w, dx, lddx, cin, mask, ldmask, mask_act, accumulate, nullptr, n, h, wd,
                          ws, ws_bytes, nullptr, stream);
}

extern "C" int b2u_conv3x3_wgrad(int dt, const void* x, int ldx, int cin, const void* dy, int lddy, int cout,
                                 float* dw, float* db, int n, int h, int wd, void* ws, size_t ws_bytes, void* stream) {
  if (dt == B2U_F16 && b2u_tensor_path_available() && b2u_tc_wgrad_ok(cin, cout, ldx, lddy))
    return b2u_tc_conv3x3_wgrad(x, ldx, cin, dy, lddy, cout, dw, db, n, h, wd, ws, ws_bytes, stream);
  return b2u_direct_conv3x3_wgrad(dt, x, ldx, cin, dy, lddy, cout, dw, db, n, h, wd, stream);
}

static int convt2x2_fwd_wp(int dt, const void* x, int ldx, int cin, const float* w, const float* bias, void* y, int ldy,
                           int cout, double* stats, int stats_sq_off, int n, int h, int wd, void* ws, size_t ws_bytes,
                           const void* wp, void* stream);

extern "C" int b2u_convt2x2_fwd(int dt, const void* x, int ldx, int cin, const float* w, const float* bias, void* y,
                                int ldy, int cout, double* stats, int stats_sq_off, int n, int h, int wd, void* ws,
                                size_t ws_bytes, void* stream) {
  return convt2x2_fwd_wp(dt, x, ldx, cin, w, bias, y, ldy, cout, stats, stats_sq_off, n, h, wd, ws, ws_bytes, nullptr,
                         stream);
}

static int convt2x2_fwd_wp(int dt, const void* x, int ldx, int cin, const float* w, const float* bias, void* y, int ldy,
                           int cout, double* stats, int stats_sq_off, int n, int h, int wd, void* ws, size_t ws_bytes,
                           const void* wp, void* stream) {
  if (dt == B2U_F16 && b2u_tensor_path_available() && b2u_tc_convt_ok(cin, cout, ldx, ldy))
    return b2u_tc_convt_fwd(x, ldx, cin, w, bias, y, ldy, cout, stats, stats_sq_off, n, h, wd, ws, ws_bytes, wp, stream);
  int rc = b2u_direct_convt_fwd(dt, x, ldx, cin, w, bias, y, ldy, cout, n, h, wd, stream);
  if (rc != B2U_OK || stats == nullptr) return rc;
  return b2u_bn_stats_off(dt, y, ldy, cout, 4LL * n * h * wd, stats, stats_sq_off, stream);   // exact path: extra pass
}

static int convt2x2_dgrad_cs(int dt, const void* dy, int lddy, int cout, const float* w, void* dx, int lddx, int cin,
                             const void* mask, int ldmask, int mask_act, int accumulate, float* colsum, int n, int h,
                             int wd, void* ws, size_t ws_bytes, const void* wp, void* stream) {
  if (dt == B2U_F16 && b2u_tensor_path_available() && b2u_tc_convt_ok(cin, cout, lddx, lddy))
    return b2u_tc_convt_dgrad(dy, lddy, cout, w, dx, lddx, cin, mask, ldmask, mask_act, accumulate, colsum, n, h, wd,
                              ws, ws_bytes, wp, stream);
  int rc = b2u_direct_convt_dgrad(dt, dy, lddy, cout, w, dx, lddx, cin, mask, ldmask, mask_act, accumulate, n, h, wd,
                                  stream);
  if (rc != B2U_OK || colsum == nullptr) return rc;
  return b2u_channel_sum(dt, dx, lddx, cin, (long long)n * h * wd, colsum, stream);       // exact path: extra pass
}

extern "C" int b2u_convt2x2_dgrad(int dt, const void* dy, int lddy, int cout, const float* w, void* dx, int lddx,
                                  int cin, const void* mask, int ldmask, int mask_act, int accumulate, int n, int h,
                                  int wd, void* ws, size_t ws_bytes, void* stream) {
  return convt2x2_dgrad_cs(dt, dy, lddy, cout, w, dx, lddx, cin, mask, ldmask, mask_act, accumulate, nullptr, n, h, wd,
                           ws, ws_bytes, nullptr, stream);
}

extern "C" int b2u_convt2x2_wgrad(int dt, const void* x, int ldx, int cin, const void* dy, int lddy, int cout,
                                  float* dw, float* db, int n, int h, int wd, void* ws, size_t ws_bytes,
                                  void* stream) {
  if (dt == B2U_F16 && b2u_tensor_path_available() && b2u_tc_convt_wgrad_ok(cin, cout, ldx, lddy))
    return b2u_tc_convt_wgrad(x, ldx, cin, dy, lddy, cout, dw, db, n, h, wd, ws, ws_bytes, stream);
  return b2u_direct_convt_wgrad(dt, x, ldx, cin, dy, lddy, cout, dw, db, n, h, wd, stream);
}

// ------------------------------------------------------------------------------------------
// op-list executor
// ------------------------------------------------------------------------------------------
static int run_one(const b2u_op& o, void* ws, size_t wsb, void* comm, void* s) {
  void* const* p = o.p;
  const int64_t* i = o.i;
  const float* f = o.f;
  const int dt = o.dt & 0xff;          // bits 8.. are executor flags (B2U_OPF_*)
#define I(k) ((int)i[k])
  switch (o.kind) {
    case B2U_OP_CONV3X3_FWD:         // p[5] (optional): packed weights
      return conv3x3_fwd_wp(dt, p[0], I(0), I(1), (const float*)p[1], (const float*)p[2], I(2), p[3], I(3), I(4),
                            (double*)p[4], I(5), I(6), I(7), ws, wsb, p[5], p[6], s,      // p[6] (optional): 1-bit ReLU mask out
                            I(8), I(9),                           // i[8]: op of a training-mode plan, i[9]: real Cin if padded
                            (const float*)p[7], (const float*)p[8]);      // optional: post-activation scale / shift
    case B2U_OP_CONV3X3_DGRAD:       // p[4] (optional): colsum
      return conv3x3_dgrad_cs(dt, p[0], I(0), I(1), (const float*)p[1], p[2], I(2), I(3), p[3], I(4), I(5), I(6),
                              (float*)p[4], I(7), I(8), I(9), ws, wsb, p[5], s);
    case B2U_OP_CONV3X3_WGRAD:
      return b2u_conv3x3_wgrad(dt, p[0], I(0), I(1), p[1], I(2), I(3), (float*)p[2], (float*)p[3], I(4), I(5), I(6), ws,
                               wsb, s);
    case B2U_OP_CONVT_FWD:           // p[5] (optional): packed weights
      return convt2x2_fwd_wp(dt, p[0], I(0), I(1), (const float*)p[1], (const float*)p[2], p[3], I(2), I(3),
                             (double*)p[4], I(7), I(4), I(5), I(6), ws, wsb, p[5], s);
    case B2U_OP_CONVT_DGRAD:         // p[4] (optional): colsum
      return convt2x2_dgrad_cs(dt, p[0], I(0), I(1), (const float*)p[1], p[2], I(2), I(3), p[3], I(4), I(5), I(6),
                               (float*)p[4], I(7), I(8), I(9), ws, wsb, p[5], s);
    case B2U_OP_CONVT_WGRAD:
      return b2u_convt2x2_wgrad(dt, p[0], I(0), I(1), p[1], I(2), I(3), (float*)p[2], (float*)p[3], I(4), I(5), I(6), ws,
                                wsb, s);
    case B2U_OP_BN_STATS:            // i[3] (optional): offset of the squares in a wider sums buffer (default: c)
      return b2u_bn_stats_off(dt, p[0], I(0), I(1), i[2], (double*)p[1], I(3) ? I(3) : I(1), s,
                              f[0], I(4), (const b2u_step_state*)p[2], p[3]);   // f[0] = rate, i[4] = op index, p[2] = step
                                                                                // state, p[3] = keep bits out: dropout(x)
    case B2U_OP_BN_FINALIZE:
      return b2u_bn_finalize((const double*)p[0], i[0], (const float*)p[1], (const float*)p[2], (float*)p[3],
                             (float*)p[4], f[0], f[1], I(1), (float*)p[5], (float*)p[6], (float*)p[7], (float*)p[8], I(2),
                             s);
    case B2U_OP_BN_APPLY:          // p[5], i[5] = split, i[6] = ld (optional): second source tensor of a split concatenate
      return b2u_bn_apply_split(dt, p[0], I(0), p[5], I(6), I(5), p[1], I(1), I(2), i[3], (const float*)p[2],
                                (const float*)p[3], (double*)p[4], I(4), s, f[0], p[6]);    // p[6]: keep bits of dropout(x)
    case B2U_OP_BN_BWD_REDUCE:
      return b2u_bn_bwd_reduce_off(dt, p[0], I(0), p[1], I(1), I(2), i[3], (const float*)p[2], (const float*)p[3],
                                   (double*)p[4], I(4) ? I(4) : I(2), s,      // i[4] (optional): offset of the second sums
                                   f[0], p[5]);                              // p[5]: keep bits of dropout(x)
    case B2U_OP_BN_BWD_APPLY:        // p[10] (optional): colsum
      return b2u_bn_bwd_apply_cs(dt, p[0], I(0), p[1], I(1), p[2], I(2), I(3), i[4], i[7], (const float*)p[3],
                                 (const float*)p[4], (const float*)p[5], (const double*)p[6], (float*)p[7],
                                 (float*)p[8], p[9], I(5), I(6), (float*)p[10], s,
                                 p[11], I(9), p[12], I(10), I(8),      // optional second (x, dx) pair, i[8] = split
                                 f[0], p[13]);                         // p[13]: keep bits of dropout(x)
    case B2U_OP_MAXPOOL_FWD:
      return b2u_maxpool_fwd(dt, p[0], I(0), p[1], I(1), I(2), I(3), I(4), I(5), f[0], I(6),
                             (const b2u_step_state*)p[2], s);
    case B2U_OP_MAXPOOL_BWD:
      return b2u_maxpool_bwd(dt, p[0], I(0), p[1], I(1), p[2], I(2), I(3), I(4), I(5), I(6), f[0], I(7),
                             (const b2u_step_state*)p[3], I(8), (double*)p[4], (const float*)p[5], (const float*)p[6], s);
    case B2U_OP_DROPOUT_FWD:
      return b2u_dropout_fwd(dt, p[0], I(0), p[1], I(1), I(2), i[3], f[0], I(4), (const b2u_step_state*)p[2], s);
    case B2U_OP_DROPOUT_BWD:
      return b2u_dropout_bwd(dt, p[0], I(0), p[1], I(1), I(2), i[3], f[0], I(4), (const b2u_step_state*)p[2], p[3], I(5),
                             I(6), s);
    case B2U_OP_COPY_SLICE:
      return b2u_copy_slice(dt, p[0], I(0), p[1], I(1), I(2), i[3], I(4), s);
    case B2U_OP_HEAD_FWD:
      return b2u_head_fwd(dt, p[0], I(0), I(1), (const float*)p[1], (const float*)p[2], (float*)p[3], i[2], s);
    case B2U_OP_BCE_DICE_SUMS:
      return b2u_bce_dice_sums((const float*)p[0], (const float*)p[1], i[0], (double*)p[2], s);
    case B2U_OP_BCE_DICE_FINALIZE:
      return b2u_bce_dice_finalize((const double*)p[0], i[0], (float*)p[1], s);
    case B2U_OP_HEAD_BWD:            // p[9] (optional): colsum
      return b2u_head_bwd_cs(dt, (const float*)p[0], (const float*)p[1], (const double*)p[2], i[0],
                             (const b2u_step_state*)p[3], p[4], I(1), I(2), (const float*)p[5], p[6], I(3), I(4),
                             (float*)p[7], (float*)p[8], i[5], (float*)p[9], s);
    case B2U_OP_DENSE_FWD:
      return b2u_dense_fwd_ws(dt, p[0], I(0), (const float*)p[1], (const float*)p[2], I(1), p[3], I(2), I(3), ws, wsb, s);
    case B2U_OP_DENSE_BWD:
      return b2u_dense_bwd(dt, p[0], I(0), (const float*)p[1], p[2], I(1), p[3], p[4], p[5], I(2), (float*)p[6],
                           (float*)p[7], I(3), I(4), s);
    case B2U_OP_BCE_FWD:
      return b2u_bce_fwd((const float*)p[0], (const float*)p[1], (const float*)p[2], I(0), (float*)p[3], s);
    case B2U_OP_BCE_SIGMOID_BWD:
      return b2u_bce_sigmoid_bwd(dt, (const float*)p[0], (const float*)p[1], (const float*)p[2], I(0),
                                 (const b2u_step_state*)p[3], p[4], s);
    case B2U_OP_ADAM:
      return b2u_adam((float*)p[0], (const float*)p[1], (float*)p[2], (float*)p[3], i[0], (b2u_step_state*)p[4], s);
    case B2U_OP_MEMSET:
      B2U_CHECK_CUDA(cudaMemsetAsync(p[0], 0, (size_t)i[0], (cudaStream_t)s));
      return B2U_OK;
    case B2U_OP_ALLREDUCE_F32:
      return b2u_allreduce(comm, p[0], i[0], 0, s);
    case B2U_OP_ALLREDUCE_F64:
      return b2u_allreduce(comm, p[0], i[0], 1, s);
    case B2U_OP_STATE_ADVANCE:
      return b2u_state_advance((b2u_step_state*)p[0], s);
    case B2U_OP_GATHER_BATCH:
      return b2u_gather_batch(dt, (const float*)p[0], (const int*)p[1], p[2], i[0], I(1), s);
    case B2U_OP_BN_APPLY_POOL:
      return b2u_bn_apply_pool(dt, p[0], I(0), p[1], I(1), I(2), I(5), I(6), I(7), (const float*)p[2], (const float*)p[3],
                               (double*)p[4], I(4), p[5], I(8), f[0], I(9), (const b2u_step_state*)p[6], s);
    case B2U_OP_BN_BWD_SUMS_WGRAD:
      return b2u_bn_bwd_sums_from_wgrad((const float*)p[0], (const float*)p[1], (const float*)p[2], (const float*)p[3],
                                        (const float*)p[4], (double*)p[5], I(0), I(1), I(2), s);
    case B2U_OP_PACK_WEIGHTS:
      return b2u_pack_weights((const long long*)p[0], I(0), (const float*)p[1], p[2], i[1], s);
    default:
      b2u_set_error("run_ops: unknown op kind %d", o.kind);
      return B2U_ERR_ARG;
  }
#undef I
}

// Side stream: ops flagged B2U_OPF_SIDE (the planner flags the weight-gradient kernels: nothing but the optimizer
// and the BN-statistics identity reads what they write) are issued on a second stream forked from the main one by an
// event, so they overlap the HBM-bound kernels of the main chain; an op flagged B2U_OPF_JOIN, and the end of the
// list, wait for everything issued there.  Works unchanged under stream capture (the side stream joins the capture
// through the fork event and is joined back before the list ends).
int g_b2u_side_stream = 0;             // b2u_set_option("side_stream", 0/1); side_init() runs when it is enabled
// ops flagged B2U_OPF_COMM (the planner's gradient-bucket all-reduces) are forked to the same side stream: a bucket's
// exchange over NVLink runs while the main stream computes the gradients of the layers below it
// (b2u_set_option("comm_overlap", 0) serialises them on the main stream for A/B runs)
int g_b2u_comm_overlap = 1;
static cudaStream_t g_side = nullptr;
static cudaEvent_t g_ev_fork = nullptr, g_ev_join = nullptr;

static int side_init() {
  if (g_side != nullptr) return B2U_OK;
  B2U_CHECK_CUDA(cudaStreamCreateWithFlags(&g_side, cudaStreamNonBlocking));
  B2U_CHECK_CUDA(cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming));
  B2U_CHECK_CUDA(cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming));
  return B2U_OK;
}

extern "C" int b2u_run_ops(const b2u_op* h_ops, int n_ops, void* ws, size_t ws_bytes, void* comm, void* stream) {
  B2U_REQUIRE(h_ops != nullptr || n_ops == 0, "run_ops: null op list");
  cudaStream_t main_s = (cudaStream_t)stream;
  bool side_pending = false;
  int rc = B2U_OK;
  for (int k = 0; k < n_ops && rc == B2U_OK; ++k) {
    const int flags = h_ops[k].dt >> 8;
    if ((g_b2u_side_stream && (flags & (B2U_OPF_SIDE >> 8))) || (g_b2u_comm_overlap && (flags & (B2U_OPF_COMM >> 8)))) {
      rc = side_init();
      if (rc != B2U_OK) break;
      B2U_CHECK_CUDA(cudaEventRecord(g_ev_fork, main_s));
      B2U_CHECK_CUDA(cudaStreamWaitEvent(g_side, g_ev_fork, 0));
      rc = run_one(h_ops[k], ws, ws_bytes, comm, (void*)g_side);
      side_pending = true;
    } else {
      if (side_pending && (flags & (B2U_OPF_JOIN >> 8))) {
        B2U_CHECK_CUDA(cudaEventRecord(g_ev_join, g_side));
        B2U_CHECK_CUDA(cudaStreamWaitEvent(main_s, g_ev_join, 0));
        side_pending = false;
      }
      rc = run_one(h_ops[k], ws, ws_bytes, comm, stream);
    }
    if (rc != B2U_OK) {
      char tmp[900];
      strncpy(tmp, g_err, sizeof(tmp) - 1);
      tmp[sizeof(tmp) - 1] = 0;
      b2u_set_error("op %d (kind %d): %s", k, h_ops[k].kind, tmp);
    }
  }
  if (side_pending) {                                    // always rejoin (also on errors: a capture must be closed)
    cudaEventRecord(g_ev_join, g_side);
    cudaStreamWaitEvent(main_s, g_ev_join, 0);
  }
  return rc;
}

// same as b2u_run_ops, with a CUDA event between consecutive ops on the launching stream:
// h_ms_out[k] = device time of op k in milliseconds (bench.py's live per-kernel roofline numbers)
extern "C" int b2u_run_ops_timed(const b2u_op* h_ops, int n_ops, void* ws, size_t ws_bytes, void* comm, void* stream,
                                 float* h_ms_out) {
  B2U_REQUIRE(h_ops != nullptr && h_ms_out != nullptr && n_ops > 0, "run_ops_timed: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  cudaEvent_t* ev = new cudaEvent_t[n_ops + 1];
  for (int k = 0; k <= n_ops; ++k) cudaEventCreate(&ev[k]);
  int rc = B2U_OK;
  cudaEventRecord(ev[0], s);
  for (int k = 0; k < n_ops && rc == B2U_OK; ++k) {
    rc = run_one(h_ops[k], ws, ws_bytes, comm, stream);
    cudaEventRecord(ev[k + 1], s);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc == B2U_OK && e == cudaSuccess) {
    for (int k = 0; k < n_ops; ++k) cudaEventElapsedTime(&h_ms_out[k], ev[k], ev[k + 1]);
  }
  for (int k = 0; k <= n_ops; ++k) cudaEventDestroy(ev[k]);
  delete[] ev;
  if (rc != B2U_OK) return rc;
  B2U_CHECK_CUDA(e);
  return B2U_OK;
}

struct b2u_graph {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
};

extern "C" int b2u_graph_create(const b2u_op* h_ops, int n_ops, void* ws, size_t ws_bytes, void* comm, void* stream,
                                void** out_graph) {
  B2U_REQUIRE(out_graph != nullptr, "graph_create: null out pointer");
  cudaStream_t s = (cudaStream_t)stream;
  B2U_REQUIRE(s != nullptr, "graph_create: capture needs a non-default stream");
  B2U_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = b2u_run_ops(h_ops, n_ops, ws, ws_bytes, comm, stream);
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(s, &g);
  if (rc != B2U_OK) {
    if (g) cudaGraphDestroy(g);
    return rc;
  }
  B2U_CHECK_CUDA(e);
  cudaGraphExec_t ex = nullptr;
  B2U_CHECK_CUDA(cudaGraphInstantiate(&ex, g, 0));
  b2u_graph* h = new b2u_graph{g, ex};
  *out_graph = h;
  return B2U_OK;
}

extern "C" int b2u_graph_launch(void* graph, void* stream) {
  B2U_REQUIRE(graph != nullptr, "graph_launch: null graph");
  B2U_CHECK_CUDA(cudaGraphLaunch(((b2u_graph*)graph)->exec, (cudaStream_t)stream));
  return B2U_OK;
}

extern "C" int b2u_graph_destroy(void* graph) {
  if (graph == nullptr) return B2U_OK;
  b2u_graph* h = (b2u_graph*)graph;
  cudaGraphExecDestroy(h->exec);
  cudaGraphDestroy(h->graph);
  delete h;
  return B2U_OK;
}

// ------------------------------------------------------------------------------------------
// NCCL (dlopen'ed so that the library loads on a box without it; torch ships libnccl.so.2)
// ------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm_t;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
  if (g_nccl.handle != nullptr) return B2U_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  const char* env = getenv("B2U_NCCL_LIB");
  if (env != nullptr) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  for (int k = 0; h == nullptr && k < 2; ++k) h = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) {
    b2u_set_error("cannot dlopen libnccl.so.2 (set B2U_NCCL_LIB): %s", dlerror());
    return B2U_ERR_NCCL;
  }
  g_nccl.GetUniqueId = (int (*)(nccl_uid*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_uid, int))dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
  g_nccl.AllReduce =
      (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
    b2u_set_error("libnccl is missing a required symbol");
    return B2U_ERR_NCCL;
  }
  g_nccl.handle = h;
  return B2U_OK;
}
#define B2U_CHECK_NCCL(expr)                                                                          \
  do {                                                                                                \
    int _r = (expr);                                                                                  \
    if (_r != 0) {                                                                                    \
      b2u_set_error("NCCL error %d: %s", _r, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
      return B2U_ERR_NCCL;                                                                            \
    }                                                                                                 \
  } while (0)
}  // namespace

extern "C" int b2u_comm_unique_id(void* h_out_128B) {
  int rc = nccl_load();
  if (rc != B2U_OK) return rc;
  nccl_uid id;
  B2U_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(h_out_128B, &id, 128);
  return B2U_OK;
}

extern "C" int b2u_comm_create(const void* h_id_128B, int rank, int world, void** out_comm) {
  int rc = nccl_load();
  if (rc != B2U_OK) return rc;
  nccl_uid id;
  memcpy(&id, h_id_128B, 128);
  nccl_comm_t c = nullptr;
  B2U_CHECK_NCCL(g_nccl.CommInitRank(&c, world, id, rank));
  *out_comm = c;
  return B2U_OK;
}

extern "C" int b2u_comm_destroy(void* comm) {
  if (comm == nullptr) return B2U_OK;
  B2U_CHECK_NCCL(g_nccl.CommDestroy((nccl_comm_t)comm));
  return B2U_OK;
}

extern "C" int b2u_allreduce(void* comm, void* buf, long long count, int is_double, void* stream) {
  B2U_REQUIRE(comm != nullptr, "allreduce: no communicator (single-GPU plans must not contain ALLREDUCE ops)");
  // ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0
  B2U_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, is_double ? 8 : 7, 0, (nccl_comm_t)comm,
                                  (cudaStream_t)stream));
  __atomic_fetch_add(&g_b2u_launches, 1ULL, __ATOMIC_RELAXED);
  return B2U_OK;
}
