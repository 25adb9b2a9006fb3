// Internal (non-ABI) entry points shared between translation units.
#pragma once
#include <stddef.h>

// exact CUDA-core path (conv_direct.cu)
int b2u_direct_conv3x3(int dt, const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act,
                       void* y, int ldy, int J, double* stats, const void* mask, int ldmask, int mask_act,
                       int accumulate, int n, int h, int wd, void* stream, void* relu_bits = nullptr,
                       int* bits_done = nullptr);     // relu_bits: written in the same pass where the kernel can (*bits_done)
int b2u_direct_conv3x3_wgrad(int dt, const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw,
                             float* db, int n, int h, int wd, void* stream);
int b2u_direct_convt_fwd(int dt, const void* x, int ldx, int cin, const float* w, const float* bias, void* y, int ldy,
                         int cout, int n, int h, int wd, void* stream);
int b2u_direct_convt_dgrad(int dt, const void* dy, int lddy, int cout, const float* w, void* dx, int lddx, int cin,
                           const void* mask, int ldmask, int mask_act, int accumulate, int n, int h, int wd,
                           void* stream);
int b2u_direct_convt_wgrad(int dt, const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw,
                           float* db, int n, int h, int wd, void* stream);

// tcgen05 tensor path (conv_tc.cu); fp16 storage, fp32 accumulation in TMEM
int b2u_tc_compiled(void);
int b2u_tc_conv3x3_ok(int k, int j, int ld_in, int ld_out);
int b2u_tc_wgrad_ok(int cin, int cout, int ldx, int lddy);
int b2u_tc_convt_ok(int cin, int cout, int ld_small, int ld_big);
int b2u_tc_convt_wgrad_ok(int cin, int cout, int ldx, int lddy);
extern int g_b2u_tc_halo;
extern int g_b2u_tc_bgroup;
extern int g_b2u_tc_max_ctas;
extern int g_b2u_tc_mcast;
extern int g_b2u_bn_async;
extern int g_b2u_wgrad_halo;
extern int g_b2u_wgrad_dhm;
extern int g_b2u_convt_jt;
extern int g_b2u_tc_2sm_max_j;
extern int g_b2u_tc_3sm;
extern long long* g_b2u_dbg;
// `wp` (optional): fp16 weights already packed by b2u_pack_weights (else packed into `ws` by the call)
// `colsum` (optional): per-channel sums of the values written, added with fp32 atomics -- the bias gradient of the
// layer whose output gradient the call produces (saves a separate pass over that gradient)
int b2u_tc_conv3x3_halo(const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act, void* y,
                        int ldy, int J, double* stats, float* colsum, const void* mask, int ldmask, int mask_act,
                        int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp, void* stream,
                        void* relu_bits_out, const float* post_scale = nullptr, const float* post_shift = nullptr);
int b2u_tc_conv3x3(const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act, void* y,
                   int ldy, int J, double* stats, float* colsum, const void* mask, int ldmask, int mask_act,
                   int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp, void* stream,
                   void* relu_bits_out);
// packed 1-bit ReLU masks (bit pix*C + c): generic producers / consumers beside the tensor-core epilogues
int b2u_relu_bits(int dt, const void* y, int ldy, int c, long long npix, void* bits, void* stream);
int b2u_apply_relu_bits(int dt, void* dx, int lddx, int c, long long npix, const void* bits, void* stream);
int b2u_channel_sum(int dt, const void* dy, int lddy, int c, long long npix, float* db, void* stream);
int b2u_bn_bwd_apply_cs(int dt, const void* dy, int lddy, const void* x, int ldx, void* dx, int lddx, int c,
                        long long npix, long long count, const float* gamma, const float* save_mean,
                        const float* save_invstd, const double* sums, float* dgamma, float* dbeta, const void* mask,
                        int ldmask, int mask_act, float* colsum, void* stream, const void* x2 = nullptr, int ldx2 = 0,
                        void* dx2 = nullptr, int lddx2 = 0, int split = 0, float p_drop = 0.f, const void* drop_bits = nullptr);
// BatchNorm over a two-input concatenate whose inputs are two dense tensors (channels [0, split) from x, the rest from x2)
int b2u_bn_apply_split(int dt, const void* x, int ldx, const void* x2, int ldx2, int split, void* y, int ldy, int c,
                       long long npix, const float* scale, const float* shift, double* out_stats, int out_sq_off,
                       void* stream, float p_drop = 0.f, const void* drop_bits = nullptr);
int b2u_tc_conv3x3_wgrad(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw, float* db,
                         int n, int h, int wd, void* ws, size_t ws_bytes, void* stream);
int b2u_tc_convt_fwd(const void* x, int ldx, int cin, const float* w, const float* bias, void* y, int ldy, int cout,
                     double* stats, int stats_sq_off, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp,
                     void* stream);
// p_drop > 0: the tensor read as the BN input is dropout(x) of a Dropout layer that is not materialised; the statistics pass
// generates the keep mask (Philox: op_id, d_state) and stores it as packed bits, the other passes read `drop_bits`
int b2u_bn_stats_off(int dt, const void* x, int ldx, int c, long long npix, double* sums, int sq_off, void* stream,
                     float p_drop = 0.f, int op_id = 0, const b2u_step_state* d_state = nullptr, void* drop_bits = nullptr);
int b2u_bn_bwd_reduce_off(int dt, const void* dy, int lddy, const void* x, int ldx, int c, long long npix,
                          const float* save_mean, const float* save_invstd, double* sums, int sq_off, void* stream,
                          float p_drop = 0.f, const void* drop_bits = nullptr);
int b2u_tc_convt_dgrad(const void* dy, int lddy, int cout, const float* w, void* dx, int lddx, int cin,
                       const void* mask, int ldmask, int mask_act, int accumulate, float* colsum, int n, int h, int wd,
                       void* ws, size_t ws_bytes, const void* wp, void* stream);
int b2u_tc_convt_wgrad(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw, float* db,
                       int n, int h, int wd, void* ws, size_t ws_bytes, void* stream);
int b2u_head_bwd_cs(int dt, const float* prob, const float* target, const double* sums, long long count,
                    const struct b2u_step_state* d_state, const void* x, int ldx, int cin, const float* w, void* dx,
                    int lddx, int x_act, float* dw, float* db, long long npix, float* colsum, void* stream);
// thin layers (Cout <= 64) with the three dw taps side by side in N (conv_tc3w.cu); b2u_set_option("tc_dwmerge", 0|1|2)
extern int g_b2u_tc_dwmerge;
extern int g_b2u_tc_dw_epi8;
int b2u_tc_conv3x3_dwmerge_ok(int K, int J);
int b2u_tc_conv3x3_dwmerge(const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act, void* y,
                           int ldy, int J, double* stats, float* colsum, const void* mask, int ldmask, int mask_act,
                           int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp, void* stream,
                           void* relu_bits_out, int packed_shift);
extern int g_b2u_tc_dw_packed;
int b2u_dense_fwd_ws(int dt, const void* x, int k, const float* w, const float* bias, int act, void* y, int m, int n,
                     void* ws, size_t ws_bytes, void* stream);
// thinnest layers (Cout = 16 / 32): row strips, dh taps merged in N, shuffle-free epilogue (conv_tc3r.cu);
// b2u_set_option("tc_rowstrip", 0|1|2)
extern int g_b2u_tc_rowstrip;
int b2u_tc_conv3x3_rowstrip_ok(int K, int J, int wd);
int b2u_tc_conv3x3_rowstrip(const void* x, int ldx, int K, const float* w, int dgrad, const float* bias, int act, void* y,
                            int ldy, int J, double* stats, float* colsum, const void* mask, int ldmask, int mask_act,
                            int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, const void* wp, void* stream,
                            void* relu_bits_out);
