/* libb200unet -- C-ABI of the B200-native U-Net train/infer hot path.
 *
 * The reference (deadskull7/One-Stop-for-COVID-19-...) has no FFI of its own: its runners call Keras
 * objects.  Every entry point below therefore cites the Keras call it replaces
 * (paths relative to /root/reference/Scripts; T1H = task1_preprocessing_plus_unet_with_comments.py,
 * UPP = task1_unet_plus_plus.py, T2 = task2_covid19_classifcation.py).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; b2u_last_error() gives the message
 *     (thread-local).  Nothing throws, nothing calls exit().
 *   - all pointers are DEVICE pointers unless the name starts with h_; the caller owns all memory
 *     (the library never allocates on the hot path); all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*) and is asynchronous.
 *   - activations are NHWC "views": base pointer + `ld` = element stride between consecutive pixels
 *     (ld >= C), so that a channel slice of a wider concat buffer is addressable in place
 *     (keras `concatenate`, T1H:887 -- zero-copy).
 *   - dt selects the storage type of activations/activation-gradients: B2U_F32 (exact mode, CUDA-core
 *     kernels) or B2U_F16 (tensor mode: tcgen05 kernels, fp32 accumulation).  Parameters, parameter
 *     gradients, optimizer state, BN statistics and the loss are always fp32/fp64.
 *   - weights stay in Keras layout: Conv2D (kh,kw,Cin,Cout); Conv2DTranspose (kh,kw,Cout,Cin);
 *     Dense (in,out).
 *   - ws / ws_bytes: scratch the op may use (b2u_ws_bytes() gives a safe upper bound).
 */
#ifndef B200UNET_H
#define B200UNET_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2U_VERSION 100

enum { B2U_OK = 0, B2U_ERR_ARG = -1, B2U_ERR_CUDA = -2, B2U_ERR_UNSUPPORTED = -3, B2U_ERR_NCCL = -4 };
enum { B2U_F32 = 0, B2U_F16 = 1 };
enum { B2U_ACT_NONE = 0, B2U_ACT_RELU = 1, B2U_ACT_ELU = 2, B2U_ACT_SIGMOID = 3,
       /* mask_act of the backward ops only: `mask` is a packed 1-bit ReLU mask (bit pix*C + c, little-endian bytes) written by
        * the forward conv (op lists: CONV3X3_FWD p[6]) instead of the fp16 activation tensor itself */
       B2U_ACT_RELU_BITS = 4 };

/* Device-resident per-step scalars, read by dropout / Adam / loss kernels so that a captured CUDA
 * graph never bakes them in (lr is set per epoch by CosineAnnealingScheduler, T1H:980-982). */
typedef struct b2u_step_state {
  uint64_t seed;        /* dropout stream key                                  */
  uint64_t step;        /* global step counter (dropout counter word 1)        */
  float lr;             /* Adam learning rate (T1H:1053 lr=0.0005)             */
  float beta1, beta2, eps;
  float beta1_pow;      /* beta1^t, beta2^t for the CURRENT step t (>=1)       */
  float beta2_pow;
  float loss_scale;     /* gradients are multiplied by this in the loss backward (fp16 range) */
  float grad_div;       /* Adam divides gradients by loss_scale*grad_div (grad_div = world size) */
  uint32_t overflow;    /* sticky: != 0 once a step was skipped because of a non-finite gradient (host resets) */
  uint32_t skip_step;   /* set by b2u_adam's check pass when the CURRENT step's gradient is non-finite: Adam then
                         * changes nothing, b2u_state_advance halves loss_scale and does not advance step / beta powers */
} b2u_step_state;

int b2u_version(void);
const char* b2u_last_error(void);
size_t b2u_ws_bytes(void);
/* 1 if the tcgen05 tensor path is compiled in and the current device is sm_100 */
int b2u_tensor_path_available(void);
/* tuning / test switches: "tc_halo" = 0 per-tap TMA loads, 1 halo tile (default), 2 halo tile with
 * descriptor base offsets, 3 three-box halo; "tensor_path" = 0/1.  Returns the previous value (<0: unknown). */
int b2u_set_option(const char* name, int value);

/* ---- step state ---------------------------------------------------------------------------- */
/* step += 1; beta_pows *= beta (device-side, graph-capturable) */
int b2u_state_advance(b2u_step_state* d_state, void* stream);

/* ---- Conv2D 3x3 'same' (T1H:859 ... :911; UPP:878; T2:748) ---------------------------------- */
/* y = act(x (*) w + bias); optionally accumulates per-channel sum / sum-of-squares of y into
 * stats[2*cout] (double) for the BatchNormalization that follows (T1H:861). */
int b2u_conv3x3_fwd(int dt, const void* x, int ldx, int cin, const float* w, const float* bias, int act,
                    void* y, int ldy, int cout, double* stats, int n, int h, int wd,
                    void* ws, size_t ws_bytes, void* stream);
/* dx = dgrad(dy, w) [* act'(mask)] ; mask is the (post-activation) tensor whose producer's
 * pre-activation gradient is wanted (NULL: none). accumulate!=0: dx += ... */
int b2u_conv3x3_dgrad(int dt, const void* dy, int lddy, int cout, const float* w,
                      void* dx, int lddx, int cin, const void* mask, int ldmask, int mask_act,
                      int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, void* stream);
/* dw += x^T (*) dy ; db += sum(dy)   (fp32, Keras layout) */
int b2u_conv3x3_wgrad(int dt, const void* x, int ldx, int cin, const void* dy, int lddy, int cout,
                      float* dw, float* db, int n, int h, int wd, void* ws, size_t ws_bytes, void* stream);

/* ---- Conv2DTranspose 2x2 stride 2 (T1H:886, 893, 900, 907; UPP:890 ...) --------------------- */
/* y[n,2i+a,2j+b,co] = sum_ci x[n,i,j,ci] * w[a,b,co,ci] + bias[co]; (h,wd) are INPUT dims */
/* stats != NULL: also accumulate per-channel sum / sum-of-squares of y for a following BatchNormalization
 * (stats[c] += sum, stats[stats_sq_off + c] += sum of squares; the BN may span a wider concat buffer) */
int b2u_convt2x2_fwd(int dt, const void* x, int ldx, int cin, const float* w, const float* bias,
                     void* y, int ldy, int cout, double* stats, int stats_sq_off, int n, int h, int wd,
                     void* ws, size_t ws_bytes, void* stream);
int b2u_convt2x2_dgrad(int dt, const void* dy, int lddy, int cout, const float* w,
                       void* dx, int lddx, int cin, const void* mask, int ldmask, int mask_act,
                       int accumulate, int n, int h, int wd, void* ws, size_t ws_bytes, void* stream);
int b2u_convt2x2_wgrad(int dt, const void* x, int ldx, int cin, const void* dy, int lddy, int cout,
                       float* dw, float* db, int n, int h, int wd, void* ws, size_t ws_bytes, void* stream);

/* ---- BatchNormalization (T1H:861, 888 ...; momentum .99, eps 1e-3) --------------------------- */
int b2u_bn_stats(int dt, const void* x, int ldx, int c, long long npix, double* sums, void* stream);
/* training: batch stats from sums -> scale/shift, save_mean/save_invstd, moving-stat update.
 * inference: scale/shift from the moving statistics. */
int b2u_bn_finalize(const double* sums, long long count, const float* gamma, const float* beta,
                    float* moving_mean, float* moving_var, float momentum, float eps, int training,
                    float* scale, float* shift, float* save_mean, float* save_invstd, int c, void* stream);
/* out_stats != NULL: also accumulate sum / sum-of-squares of the values written to y (the statistics a
 * BatchNormalization over a concat buffer that contains y needs): out_stats[c], out_stats[out_sq_off + c] */
int b2u_bn_apply(int dt, const void* x, int ldx, void* y, int ldy, int c, long long npix,
                 const float* scale, const float* shift, double* out_stats, int out_sq_off, void* stream);
/* sums[0:c] += sum(dy), sums[c:2c] += sum(dy * xhat) */
int b2u_bn_bwd_reduce(int dt, const void* dy, int lddy, const void* x, int ldx, int c, long long npix,
                      const float* save_mean, const float* save_invstd, double* sums, void* stream);
/* dx = gamma*invstd*(dy - sum_dy/count - xhat*sum_dyxhat/count) [* act'(mask)]; dgamma += ..; dbeta += ..
 * (count == npix on one GPU; the global pixel count when the sums were all-reduced) */
int b2u_bn_bwd_apply(int dt, const void* dy, int lddy, const void* x, int ldx, void* dx, int lddx,
                     int c, long long npix, long long count, const float* gamma, const float* save_mean,
                     const float* save_invstd, const double* sums, float* dgamma, float* dbeta,
                     const void* mask, int ldmask, int mask_act, void* stream);

/* ---- MaxPooling2D((2,2)) [+ Dropout(p)] (T1H:862-863) ---------------------------------------- */
/* (h,wd) are INPUT dims. p_drop==0 -> plain pooling. */
int b2u_maxpool_fwd(int dt, const void* x, int ldx, void* y, int ldy, int c, int n, int h, int wd,
                    float p_drop, int op_id, const b2u_step_state* d_state, void* stream);
/* dx[first arg-max of each window] (+)= dy*keep/(1-p), other window elements (+)= 0 */
int b2u_maxpool_bwd(int dt, const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int c,
                    int n, int h, int wd, float p_drop, int op_id, const b2u_step_state* d_state,
                    int accumulate, double* bn_sums, const float* bn_gamma, const float* bn_beta, void* stream);
/* bn_sums != NULL: x is the output of a training-mode BatchNormalization (x = gamma*xhat + beta) and this call
 * completes its gradient; the kernel then also accumulates that BN's backward statistics from the values it
 * already holds -- bn_sums[0:c] += sum(dx), bn_sums[c:2c] += sum(dx * xhat), xhat = (x - beta)/gamma -- which
 * replaces a separate b2u_bn_bwd_reduce pass over dx and the BN input. */
/* ---- Dropout(p) (UPP:864, 879; T2:777) -------------------------------------------------------- */
int b2u_dropout_fwd(int dt, const void* x, int ldx, void* y, int ldy, int c, long long npix, float p,
                    int op_id, const b2u_step_state* d_state, void* stream);
int b2u_dropout_bwd(int dt, const void* dy, int lddy, void* dx, int lddx, int c, long long npix, float p,
                    int op_id, const b2u_step_state* d_state, const void* mask, int ldmask, int mask_act,
                    void* stream);
/* ---- concatenate helpers (UPP:905: one tensor feeding several concats) ------------------------ */
int b2u_copy_slice(int dt, const void* src, int ldsrc, void* dst, int lddst, int c, long long npix,
                   int accumulate, void* stream);

/* ---- output head: Conv2D(1,(1,1),activation='sigmoid') (T1H:913) + bce_dice_loss (T1H:784-799) - */
int b2u_head_fwd(int dt, const void* x, int ldx, int cin, const float* w, const float* bias,
                 float* prob, long long npix, void* stream);
/* sums[0..3] += sum(t*p), sum(t), sum(p), sum(bce) */
int b2u_bce_dice_sums(const float* prob, const float* target, long long count, double* sums, void* stream);
/* out[0] = 0.5*bce_mean + 0.5*(1-dice), out[1] = dice */
int b2u_bce_dice_finalize(const double* sums, long long count, float* out, void* stream);
/* dlogit = loss_scale * dL/dp * p(1-p); dx = dlogit * w [* act'(x)]; dw += sum dlogit*x; db += sum dlogit.
 * global_count/global sums allow a data-parallel batch to use whole-batch Dice statistics. */
int b2u_head_bwd(int dt, const float* prob, const float* target, const double* sums, long long count,
                 const b2u_step_state* d_state, const void* x, int ldx, int cin, const float* w,
                 void* dx, int lddx, int x_act, float* dw, float* db, long long npix, void* stream);

/* ---- Dense (T2:776-778) ------------------------------------------------------------------------ */
/* x (n,k) of storage type dt -> y (n,m) ALWAYS fp32 (dense outputs / their gradients are tiny) */
int b2u_dense_fwd(int dt, const void* x, int k, const float* w, const float* bias, int act,
                  void* y, int m, int n, void* stream);
int b2u_dense_bwd(int dt, const void* x, int k, const float* w, const void* y, int act, const void* dy,
                  void* dx, const void* mask, int mask_act, float* dw, float* db, int m, int n, void* stream);
/* weighted binary cross-entropy on (n,1) probabilities (T2:828, class_weight T2:836) */
int b2u_bce_fwd(const float* prob, const float* target, const float* sample_w, int n, float* out, void* stream);
/* dlogit for a sigmoid unit, written as storage type dt: dy[i] = loss_scale * w_i (p-t)/n (clip-aware) */
int b2u_bce_sigmoid_bwd(int dt, const float* prob, const float* target, const float* sample_w, int n,
                        const b2u_step_state* d_state, void* dlogit, void* stream);

/* ---- Adam (T1H:1053) --------------------------------------------------------------------------- */
int b2u_adam(float* params, const float* grads, float* m, float* v, long long n,
             b2u_step_state* d_state, void* stream);

/* ---- host-array ingest: fit/predict batches (T1H:1059 model.fit(x_train, ...)) ----------------- */
/* dst[b, :] = (T) src[idx[b], :]  (src fp32, device-resident dataset; idx int32 device) */
int b2u_gather_batch(int dt, const float* src, const int* idx, void* dst, long long per_sample, int nb,
                     void* stream);
/* the same with the channel count zero-padded from c to cpad: inference plans feed a first conv of 2..15 input channels
 * (T2:748 on 224 x 224 x 3 slices) to the tensor-core kernel through a 16-channel input tensor */
int b2u_gather_batch_pad(int dt, const float* src, const int* idx, void* dst, long long pix_per_sample, int c, int cpad,
                         int nb, void* stream);

/* ---- sm.metrics threshold sweep (T1H:1206-1211) ------------------------------------------------ */
/* for each threshold k: tp[k] += sum(t * (p > thr[k])), sum_pr[k] += sum(p > thr[k]); sum_gt += sum(t) */
int b2u_threshold_counts(const float* prob, const float* target, long long count, const float* thresholds,
                         int nthr, double* tp, double* sum_pr, double* sum_gt, void* stream);

/* ---- preprocessing (T1H:163-194 clahe_enhancer, T1H:211-273 cropper, T1H:485-488 resize) ------- */
/* cv2.createCLAHE(clipLimit, (tiles,tiles)).apply on n uint8 images (h % tiles == 0, w % tiles == 0) */
int b2u_clahe_u8(const uint8_t* in, uint8_t* out, int n, int h, int wd, float clip_limit, int tiles,
                 void* ws, size_t ws_bytes, void* stream);
/* per image: crop two boxes (x,y,w,h int32 x8), cv2.resize INTER_AREA to (half_w x out_h) each, hconcat,
 * cv2.resize INTER_LINEAR (u8 fixed point) to final x final, /255 -> float32 (n,final,final); both stages are
 * bit-exact against OpenCV (mid_u8 = the 250 x 250 stage, out = np.uint8(.)/255) */
int b2u_crop_resize(const uint8_t* in, int n, int h, int wd, const int* boxes, int half_w, int out_h,
                    int final_dim, uint8_t* mid_u8, float* out, void* stream);
/* cv2.resize(img, (dst_w, dst_h), interpolation) on n uint8 images (T1H:335 the 512 x 512 INTER_AREA stage of the
 * NIfTI ingest; T1H:485-488 INTER_LINEAR to new_dim): interpolation = OpenCV's enum, 1 INTER_LINEAR (square targets),
 * 3 INTER_AREA.  Bit-exact against OpenCV's CV_8UC1 arithmetic (area tables / 11-bit fixed point). */
int b2u_resize_u8(const uint8_t* in, int n, int src_h, int src_w, uint8_t* out, int dst_h, int dst_w,
                  int interpolation, void* stream);
/* read_nii's slice stage (T1H:288-297, 335-337): cv2.resize(slice, (dst_w, dst_h), INTER_AREA) on n float64 slices
 * (get_fdata() values), bit-exact against OpenCV's CV_64F path, then -- minmax_normalize != 0 -- (img - min) / (max - min)
 * per slice in double (a constant slice becomes NaN like the reference's 0/0) */
int b2u_resize_area_f64(const double* in, int n, int src_h, int src_w, double* out, int dst_h, int dst_w,
                        int minmax_normalize, void* stream);

/* ---- BatchNormalization apply + MaxPooling2D((2,2)) + Dropout in one pass (encoder level, T1H:861-863) ---- */
/* y = x * scale + shift is written at full resolution (the skip tensor, possibly a concat slice) and its 2x2 max,
 * taken over the values as stored and passed through the dropout of b2u_maxpool_fwd, at half resolution into yp --
 * the skip tensor is not read back.  out_stats as in b2u_bn_apply (statistics of y for a BN over the concat buffer).
 * Op record: p = {x, y, scale, shift, out_stats, yp, d_state}, i = {ldx, ldy, c, n*h*wd, out_sq_off, n, h, wd, ldp,
 * op_id}, f = {p_drop}. */
int b2u_bn_apply_pool(int dt, const void* x, int ldx, void* y, int ldy, int c, int n, int h, int wd, const float* scale,
                      const float* shift, double* out_stats, int out_sq_off, void* yp, int ldp, float p_drop, int op_id,
                      const b2u_step_state* d_state, void* stream);

/* ---- BatchNormalization backward statistics without a pass over the activations ------------------ */
/* For a BN whose output y feeds exactly one Conv2D 3x3 (U-Net decoder: concat -> BN -> conv, T1H:888-889) the two
 * per-channel sums its backward needs follow from quantities the conv backward already produced:
 *   sum_p dy[p][c] * y[p][c] = sum_{tap,co} W[tap][c][co] * dW[tap][c][co]      (adjoint identity of the convolution)
 *   sum_p dy[p][c]           = `colsum` emitted by the conv's data-gradient kernel
 * so sums[c] = colsum[c] and sums[C + c] = sum dy * xhat = (<W, dW>_c - beta[c] * colsum[c]) / gamma[c].
 * w, dw: Keras HWIO (taps, C, cout) fp32; dw must hold only this step's gradient of that conv. */
int b2u_bn_bwd_sums_from_wgrad(const float* w, const float* dw, const float* colsum, const float* gamma,
                               const float* beta, double* sums, int c, int cout, int taps, void* stream);

/* ---- fp16 operand copies of all conv / transposed-conv kernels of a model in ONE launch ---------- */
/* The tcgen05 kernels read weights as fp16 [tap][out][in] tiles.  `d_table` holds n_entries (<= 128) records of 8
 * int64: {src element offset in params, dst element offset in wpack, first work tile, mode, taps, J, K, 0}; a work
 * tile is a 32 x 32 piece of one tap matrix, an entry has taps * ceil(J/32) * ceil(K/32) of them and `total` is their
 * sum over all entries;
 * mode 0: conv fwd  Wp[t][co][ci] = w[t][ci][co]        1: conv dgrad  Wp[t][ci][co] = w[8-t][ci][co]
 *      2: convT fwd Wp[q][ci]     = w[q][ci]             3: convT dgrad Wp[ab][ci][co] = w[ab][co][ci]
 * (Keras layouts, T1H:859 / T1H:886).  Op lists pass the packed tile address as the optional trailing pointer p[5] of
 * CONV3X3_FWD / CONV3X3_DGRAD / CONVT_FWD / CONVT_DGRAD; without it every conv call packs its own weights first. */
int b2u_pack_weights(const long long* d_table, int n_entries, const float* params, void* wpack, long long total,
                     void* stream);

/* ---- plan executor: a whole forward / train step as one array of op records ------------------- */
enum {
  B2U_OP_CONV3X3_FWD = 1, B2U_OP_CONV3X3_DGRAD, B2U_OP_CONV3X3_WGRAD,
  B2U_OP_CONVT_FWD, B2U_OP_CONVT_DGRAD, B2U_OP_CONVT_WGRAD,
  B2U_OP_BN_STATS, B2U_OP_BN_FINALIZE, B2U_OP_BN_APPLY, B2U_OP_BN_BWD_REDUCE, B2U_OP_BN_BWD_APPLY,
  B2U_OP_MAXPOOL_FWD, B2U_OP_MAXPOOL_BWD, B2U_OP_DROPOUT_FWD, B2U_OP_DROPOUT_BWD, B2U_OP_COPY_SLICE,
  B2U_OP_HEAD_FWD, B2U_OP_BCE_DICE_SUMS, B2U_OP_BCE_DICE_FINALIZE, B2U_OP_HEAD_BWD,
  B2U_OP_DENSE_FWD, B2U_OP_DENSE_BWD, B2U_OP_BCE_FWD, B2U_OP_BCE_SIGMOID_BWD,
  B2U_OP_ADAM, B2U_OP_MEMSET, B2U_OP_ALLREDUCE_F32, B2U_OP_ALLREDUCE_F64, B2U_OP_STATE_ADVANCE,
  B2U_OP_GATHER_BATCH, B2U_OP_PACK_WEIGHTS, B2U_OP_BN_BWD_SUMS_WGRAD, B2U_OP_BN_APPLY_POOL
};
/* one record; the meaning of p[]/i[]/f[] per kind is the argument order of the function above
 * (pointers in order into p[], ints/long longs into i[], floats into f[]).
 * Op lists only -- one optional trailing pointer `colsum` (fp32 [C], may be NULL) on the ops that write the final
 * gradient of a conv output: CONV3X3_DGRAD p[4], CONVT_DGRAD p[4], BN_BWD_APPLY p[10], HEAD_BWD p[9].  The kernel adds
 * the per-channel sums of the dx values it writes, which is that conv's bias gradient (Keras: the `bias` slot of
 * Conv2D, T1H:859); the matching *_WGRAD op is then given db = NULL and skips its own pass over the gradient. */
/* Further optional op-list slots (all may be NULL / 0; they extend, never reorder, the function arguments):
 *  CONV3X3_FWD   p[6] packed 1-bit ReLU mask out; p[7], p[8] per-channel scale / shift applied AFTER the activation (the
 *                inference-mode BatchNormalization behind the conv, T2:749-750, folded into the epilogue); i[8] = 1: op of a
 *                training plan; i[9]: real Cin when the input tensor is zero-padded to 16 channels.
 *  BN_STATS      i[3] offset of the squares in a wider sums buffer; f[0], i[4], p[2] (step state), p[3]: the input is
 *                dropout(x) of a Dropout layer that is not materialised (UPP:874-876) -- the kernel draws the keep mask
 *                (rate f[0], dropout op index i[4]) and stores it as packed bits (bit pix * C + c) at p[3].
 *  BN_APPLY      p[5], i[5] = split, i[6] = ld: channels [split, C) come from a second tensor (a two-input concatenate kept
 *                as two dense tensors); f[0], p[6]: dropout(x) as above, keep bits READ from p[6].
 *  BN_BWD_REDUCE i[4] offset of the second sums; f[0], p[5]: dropout(x), keep bits read from p[5].
 *  BN_BWD_APPLY  p[11], i[9] = ld, p[12], i[10] = ld, i[8] = split: second (input, input-gradient) pair of a split
 *                concatenate; f[0], p[13]: dropout(x), keep bits read from p[13] (the gradient written is the one of x). */
/* executor flags, OR-ed into b2u_op.dt above the storage type (dt & 0xff): */
#define B2U_OPF_SIDE 0x100 /* may run on the executor's side stream (forked / joined with events; weight gradients) */
#define B2U_OPF_JOIN 0x200 /* reads what earlier B2U_OPF_SIDE ops wrote: wait for the side stream first            */
#define B2U_OPF_COMM 0x400 /* gradient-bucket all-reduce: ALWAYS on the side stream (overlaps the rest of the backward) */
typedef struct b2u_op {
  int32_t kind;
  int32_t dt;
  void* p[14];
  int64_t i[12];
  float f[4];
} b2u_op;
int b2u_run_ops(const b2u_op* h_ops, int n_ops, void* ws, size_t ws_bytes, void* comm, void* stream);
/* b2u_run_ops with a CUDA event between consecutive ops; h_ms_out[k] = device milliseconds of op k.
 * Synchronises the stream before returning (profiling aid for bench.py's roofline numbers). */
int b2u_run_ops_timed(const b2u_op* h_ops, int n_ops, void* ws, size_t ws_bytes, void* comm, void* stream,
                      float* h_ms_out);
/* CUDA-graph capture of an op list (launch-bound inner loop -> one graph launch per step) */
int b2u_graph_create(const b2u_op* h_ops, int n_ops, void* ws, size_t ws_bytes, void* comm, void* stream,
                     void** out_graph);
int b2u_graph_launch(void* graph, void* stream);
int b2u_graph_destroy(void* graph);
/* number of kernels launched by this library since process start (bench.py's gpu_launches) */
long long b2u_launch_count(void);

/* ---- data-parallel gradient exchange (no reference counterpart: the reference is single-device) - */
int b2u_comm_unique_id(void* h_out_128B);
int b2u_comm_create(const void* h_id_128B, int rank, int world, void** out_comm);
int b2u_comm_destroy(void* comm);
int b2u_allreduce(void* comm, void* buf, long long count, int is_double, void* stream);

#ifdef __cplusplus
}
#endif
#endif
