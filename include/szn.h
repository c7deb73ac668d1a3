/* szn.h — C ABI of libszn.so, the sm_100a implementation of the SZN pixel-embedding hot path.
 *
 * The reference (RohanDoshi2018/ZeroshotSemanticSegmentation) has no FFI/plugin boundary: its hot path
 * is Python calling torch ops.  This header is the boundary a maintainer binds instead (ctypes stub
 * in INTEGRATION.md); each entry point names the reference lines whose device work it replaces.
 *
 * Conventions
 *   - every function returns 0 or a negative SZN_ERR_* code; szn_last_error() gives the message.
 *   - the caller owns every buffer (device pointers unless stated), nothing is allocated here,
 *     nothing synchronises; work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - dtype: SZN_F32 = fp32 storage, TF32 tensor-core products, fp32 accumulate;
 *            SZN_BF16 = bf16 storage, bf16 products, fp32 accumulate;
 *            SZN_F32X3 = fp32-grade: every value v is stored as two bf16 planes hi = bf16(v), lo = bf16(v - hi)
 *            (a pixel row of C channels is [hi_0..hi_{C-1} | lo_0..lo_{C-1}], the same 4C bytes as fp32; packed
 *            weights are two planes, hi then lo) and every product is the error-compensated sum
 *            hi*hi + lo*hi + hi*lo of three kind::f16 MMAs with fp32 accumulate (relative error ~2^-16 per product
 *            against 2^-11 for TF32): the precision the reference's fp32 CPU path is held to (models.py:114-160).
 *   - trunk activations are NHWC ([B][H][W][C], channels contiguous); weights for the tensor-core
 *     kernels are [Cout][R*S][Cin] ("OHWI") in the activation dtype; the public tensors of the
 *     reference API (input image, returned score) stay NCHW fp32.
 */
#ifndef SZN_H_
#define SZN_H_

#ifdef __cplusplus
extern "C" {
#endif

#define SZN_F32 0
#define SZN_BF16 1
#define SZN_F32X3 2

#define SZN_ERR_ARG (-1)
#define SZN_ERR_CUDA (-2)
#define SZN_ERR_UNSUPPORTED (-3)

const char* szn_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long szn_launch_count(void);
int szn_abi_version(void);

/* ---- tensor-core implicit-GEMM convolutions (models.py:45-93 conv1_2..score_fr and their autograd backward) ---- */
/* y[B,Ho,Wo,*] = act(conv(x[B,H,W,Cin], wt[Cout][R*S][Cin]) + bias) * scale[b][co];  ldo = row stride of y (elements).
 * 1x1/pad 0 convolutions are run as one flat GEMM over all B*H*W pixels. out_fp32: store raw fp32 regardless of dtype. */
int szn_conv_fwd(int dtype, const void* x, const void* wt, const float* bias, void* y, int B, int H, int W, int Cin,
                 int Cout, int R, int S, int pad, int relu, const float* scale, int scale_ld, int out_fp32,
                 long long ldo, void* stream);
/* dx[B,H,W,Cin] = conv_transpose(dy[B,Ho,Wo,Cout] (row stride ld_dy), w) * scale[b][ci] * (relu_ref[B,H,W,Cin] > 0)
 * wt_dgrad: the transposed, tap-flipped weights [Cin][R*S][Cout] written by szn_pack_weight_dgrad(mode 0); the data
 * gradient then runs as a forward conv of dy with padding R-1-pad on the same tensor-core path.
 * dx_col_sum (nullable, fp32 [Cin], zeroed by the caller): += column sums of the stored dx, i.e. the bias gradient of the
 * layer that produced this conv's input, fused into the epilogue (one less pass over dx). */
int szn_conv_dgrad(int dtype, const void* dy, const void* wt_dgrad, void* dx, int B, int H, int W, int Cin, int Cout,
                   int R, int S, int pad, const void* relu_ref, const float* scale, int scale_ld, long long ld_dy,
                   float* dx_col_sum, void* stream);
/* dw[Cout][R*S*Cin] += sum_pixels dy (x) x   (fp32; split-K partial tiles are added with TMA reduce-add: zero dw first).
 * Cin, Cout and ld_dy must be multiples of 128 bytes of the element type (32 fp32 / 64 bf16).  The buffer is the
 * channels_last image of the OIHW gradient, so it can be handed to autograd as weight.grad of a channels_last parameter
 * without a transposing pass (trainer_fcn.py:157 loss.backward()). */
int szn_conv_wgrad(int dtype, const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int Cout, int R,
                   int S, int pad, long long ld_dy, void* stream);

/* split-K work items per persistent CTA of szn_conv_wgrad (default 1; data-parallel runs use 3 so that SMs shared with the
 * gradient all-reduce kernels do not set the kernel's duration -- the tiles are handed out by a global work counter) */
int szn_set_wgrad_waves(int waves);

/* ---- conv1_1 (models.py:43-44,116): Cin=3, pad=100, CUDA cores; x is the public NCHW fp32 image ---- */
int szn_conv1_1_fwd(int dtype, const float* x, const float* w_oihw, const float* bias, void* y, int B, int H, int W,
                    int pad, void* stream);
/* dw_oihw[64][3][3][3] += ...   (fp32 atomics: zero first) */
int szn_conv1_1_wgrad(int dtype, const float* x, const void* dy, float* dw_oihw, int B, int H, int W, int pad,
                      void* stream);

/* ---- MaxPool2d(2, stride=2, ceil_mode=True) (models.py:47,54,63,72,81) on NHWC ---- */
int szn_pool_fwd(int dtype, const void* in, void* out, int B, int H, int W, int C, void* stream);
/* dy = route dp to the first maximum of each window; relu_gate additionally zeroes positions where y <= 0.
 * dy_col_sum (nullable, fp32 [C], zeroed by the caller): += per-channel sums of dy (the producer conv's bias gradient) */
int szn_pool_bwd(int dtype, const void* y, const void* dp, void* dy, int B, int H, int W, int C, int relu_gate,
                 float* dy_col_sum, void* stream);

/* The same pair with a one-byte routing code per pooled element in place of the second read of the pre-pool tensor
 * (the engine's default): code[B][Ho][Wo][C] = winner 0..3 (scan order, first maximum, as ATen's max_pool2d backward
 * picks it) | 4 when that maximum is > 0 (the ReLU gate of the producer conv, models.py:46-47 etc.).  szn_pool_bwd_code
 * gives exactly the dy of szn_pool_bwd from dp and the code alone, so the pre-pool activation is neither re-read nor kept.
 * code must be 8-byte aligned. */
int szn_pool_fwd_code(int dtype, const void* in, void* out, unsigned char* code, int B, int H, int W, int C, void* stream);
int szn_pool_bwd_code(int dtype, const unsigned char* code, const void* dp, void* dy, int B, int H, int W, int C,
                      int relu_gate, float* dy_col_sum, void* stream);

/* db[C] += column sums of dy[rows][ld]  (stand-alone bias gradient; the conv layers get theirs fused into
 * szn_conv_dgrad / szn_pool_bwd, this entry point serves the 17x17 score heads; zero db first) */
int szn_bias_grad(int dtype, const void* dy, float* db, long long rows, int C, long long ld, void* stream);

/* parameter layout conversion between the reference's OIHW fp32 tensors (state_dict layout, models.py:43-98)
 * and the kernels' [O_pad][R*S][I] layout (rows >= O zero-filled) */
int szn_pack_weight(int dtype, const float* w_oihw, void* out, int O, int I, int R, int S, int O_pad, void* stream);
/* OIHW fp32 -> [Cin][R*S flipped][O_pad] (mode 0) or [R*S*Cin][O_pad] (mode 1): K-major operands of the data gradient */
int szn_pack_weight_dgrad(int dtype, const float* w_oihw, void* out, int O, int I, int R, int S, int O_pad, int mode,
                          void* stream);
/* fc6 data gradient (models.py:84, 7x7 valid conv on a 23x23 map): dcol[B,Ho,Wo,R*S*C] (one GEMM against the mode-1
 * weights) is folded back to dx[B,H,W,C] by summing the <= R*S window positions that cover each pixel */
int szn_col2im(int dtype, const void* dcol, void* dx, int B, int H, int W, int C, int R, int S, void* stream);
int szn_unpack_wgrad(const float* dw_ohwi, float* g_oihw, int O, int I, int R, int S, void* stream);
/* fp32 [rows][C] -> the storage format `dtype` (row length C matters for SZN_F32X3's plane layout) */
int szn_cast(int dtype, const float* in, void* out, long long rows, int C, void* stream);

/* Dropout2d(p=0.5) (models.py:86,91): scale[n] in {0,2}, one value per (image, channel) */
int szn_dropout_scale(float* scale, int n, unsigned long long seed, void* stream);

/* ---- upscore / seenmask_upscore: ConvTranspose2d(k=64, stride=32) + crop 19 (models.py:94,98,146-151) ----
 * s: fp32 [B,hs,ws,ld], channels [coff, coff+D).  out / g: NCHW fp32 [B,D,H,W] (the tensors forward() returns). */
int szn_upsample32_crop_fwd(const float* s, float* out, int B, int D, int H, int W, int hs, int ws, int ld, int coff,
                            void* stream);
int szn_upsample32_crop_bwd(int dtype, const float* g, void* ds, int B, int D, int H, int W, int hs, int ws, int ld,
                            int coff, void* stream);
/* dense variant for small channel counts (trained 2x2 seenmask head); wd: [Ci][Co][64][64] fp32 */
int szn_deconv_small_fwd(const float* s, const float* wd, float* out, int B, int Ci, int Co, int H, int W, int hs,
                         int ws, int ld, int coff, void* stream);
int szn_deconv_small_dgrad(int dtype, const float* g, const float* wd, void* ds, int B, int Ci, int Co, int H, int W,
                           int hs, int ws, int ld, int coff, void* stream);
int szn_deconv_small_wgrad(const float* s, const float* g, float* dwd, int B, int Ci, int Co, int H, int W, int hs,
                           int ws, int ld, int coff, void* stream);

/* ---- losses (utils.py:19-102).  score NCHW fp32, target int64 [n,h,w] with -1 = ignore.
 * kind 0 = cosine_loss (utils.py:75-102), 1 = mse_loss (utils.py:50-73).  accum = {sum, n_valid} (fp64, device).
 * table [table_rows][c] (or null with an explicit target_embed): a label >= table_rows never reads past the table; it
 * turns the loss into NaN (torch's embedding would device-assert), likewise a label >= c in szn_ce2d_*. */
int szn_embed_loss_fwd(int kind, const float* score, const long long* target, const float* target_embed,
                       const float* table, int table_rows, int n, int c, int h, int w, float* stats, double* accum,
                       float* loss, void* stream);
int szn_embed_loss_bwd(int kind, const float* score, const long long* target, const float* target_embed,
                       const float* table, int table_rows, int n, int c, int h, int w, const float* stats,
                       const double* accum, const float* grad_out, float* dscore, void* stream);
/* cross_entropy2d (utils.py:19-48) */
int szn_ce2d_fwd(const float* score, const long long* target, int n, int c, int h, int w, int size_average, float* lse,
                 double* accum, float* loss, void* stream);
int szn_ce2d_bwd(const float* score, const long long* target, int n, int c, int h, int w, int size_average,
                 const float* lse, const double* accum, const float* grad_out, float* dscore, void* stream);
/* loss from (possibly all-reduced) accumulators: kind 0 (N-sum)/N, 1 sum/N, 2 sum, 3 sum/N */
int szn_loss_finalize(int kind, const double* accum, float* loss, void* stream);

/* ---- inference (utils.py:159-205) ---- */
/* infer_lbl: labels = argmax_c <s_p,e_c>/(|s_p| |e_c|), |e_c|==0 -> 1, lowest index on ties.
 * Runs as an error-compensated (3xTF32) tcgen05 GEMM when h*w % 32 == 0 and C <= 256, else on CUDA cores (fp32 FMA).
 * en_scratch: szn_embed_argmax_scratch_floats(C, D) floats of device memory. */
long long szn_embed_argmax_scratch_floats(int C, int D);
int szn_embed_argmax(const float* score, const float* table, int n, int D, int h, int w, int C, float* en_scratch,
                     long long* labels, void* stream);
/* stich_seen_unseen_with_mask: mask from seen_mask_score [n,2,h,w] (infer_lbl_szn) or, when that is null,
 * from target in unseen[] (infer_lbl_forced_unseen) */
int szn_stitch_labels(const long long* lbl_seen, const long long* lbl_unseen, const float* seen_mask_score,
                      const long long* target, const long long* unseen, int n_unseen, int n, int h, int w,
                      long long* out, void* stream);

/* ---- EXPERIMENTAL, opt-in (FCN32s(fused_head=True)), not on the default path: cosine loss (kind 0, utils.py:75-102) or
 * MSE loss (kind 1, utils.py:50-73), labels (utils.py:159-185) and the gradient of the hs x ws score map computed from that map alone; the (B, D, H, W) score
 * tensor (models.py:94,146-147) is neither read nor written.  s17: fp32 [B,hs,ws,ld], channels [coff, coff+D);
 * target: int64 [B,H,W] (-1 = ignore) or null (labels only); labels: int64 [B,H,W] or null; accum = {sum cos, n_valid}.
 * workspace: szn_head_fused_workspace_floats(B, hs, ws, C) floats, shared by fwd and bwd of one step.
 * bwd writes ds17 fp32 [B,hs,ws,ld]: grad_out/N * dL/ds in channels [coff, coff+D), 0 elsewhere; N = accum[1], which the
 * caller may all-reduce between the two calls (then szn_loss_finalize(kind, ...) gives the global loss). */
long long szn_head_fused_workspace_floats(int B, int hs, int ws, int C);
int szn_head_fused_fwd(int kind, const float* s17, int ld, int coff, const long long* target, const float* table, int B,
                       int D, int H, int W, int hs, int ws, int C, float* workspace, double* accum, float* loss,
                       long long* labels, void* stream);
int szn_head_fused_bwd(int kind, const float* s17, int ld, int coff, const float* table, int B, int D, int hs, int ws, int C,
                       const float* workspace, const double* accum, const float* grad_out, float* ds17, void* stream);

/* ---- optimizer step (train.py:126-129 SGD param groups; trainer_fcn.py:158 optim.step()) ----
 * One fused pass over a parameter's storage: d = g + wd*p; buf = first ? d : momentum*buf + d; p -= lr*buf.
 * param / grad / momentum_buf: n fp32 values in the SAME memory order (all three dense, same strides). */
int szn_sgd_step(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                 float weight_decay, int first_step, void* stream);

/* Adam (train.py:130-133 optional FCN optimizer, train.py:175 seen-mask phase: torch.optim.Adam(params, lr)), amsgrad off.
 * g' = g + wd*p; m += (1-b1)(g'-m); v = b2*v + (1-b2)g'^2; p -= step_size * m / (sqrt(v)/bias_correction2_sqrt + eps),
 * with step_size = lr / (1 - b1^t) and bias_correction2_sqrt = sqrt(1 - b2^t) formed by the caller for step t >= 1.
 * param / grad / exp_avg / exp_avg_sq: n fp32 values in the SAME memory order, 16-byte aligned. */
int szn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float beta1, float beta2,
                  float eps, float weight_decay, float step_size, float bias_correction2_sqrt, void* stream);

/* ---- metrics (utils.py:104-154, called after every iteration: trainer_fcn.py:164,223,248) ----
 * Confusion matrices of _fast_hist for target = 'all' (and 'seen' / 'unseen' when is_unseen[n_class] is given), built on
 * the device-resident label maps.  hist: int64 [1 or 3][n_class][n_class], accumulated into (zero it first). */
int szn_confusion_hist(const long long* label_true, const long long* label_pred, long long n, int n_class,
                       const unsigned char* is_unseen, long long* hist, void* stream);

/* ---- data parallelism: the one exchange step of the path, all-reduce(sum) of parameter gradients.  The
 * reference has no distributed code; this is what a maintainer binds for N > 1.  The communicator is NCCL's, created from a
 * 128-byte unique id that rank 0 obtains and the caller ships to every rank (ddp.py uses torch.distributed for that).  NCCL
 * is resolved at run time from the libnccl.so.2 already in the process; szn_comm_available() says whether it was found. */
int szn_comm_available(void);
int szn_comm_unique_id(char* id128);
int szn_comm_init(const char* id128, int rank, int world, void** comm_out);
int szn_comm_destroy(void* comm);
/* one bucket: n device buffers (bufs[i], counts[i] elements of `dtype`: SZN_F32, SZN_BF16 or 3 = fp64) summed over the
 * ranks in place, inside one NCCL group (one fused launch) on `stream`.  bufs / counts are HOST arrays. */
int szn_allreduce_bucket(void* comm, void* const* bufs, const long long* counts, int n, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SZN_H_ */
