/* libpcrl_b200.so -- C ABI of the B200-native PCRLv2 3-D pre-training hot path.
 *
 * The reference (RL4M/PCRLv2) is pure Python on top of PyTorch/cuDNN: it has no FFI of its own,
 * so each entry point below names the torch call of the reference it replaces
 * (paths relative to the reference root).  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (the library never allocates or
 *     retains memory), sizes are plain ints, `stream` is a cudaStream_t passed as void*;
 *   - calls are asynchronous on `stream`, never synchronise, and return 0 or a negative code
 *     (PCRL_ERR_*); pcrl_last_error() returns the message of the calling thread's last failure;
 *   - `dtype` (PCRL_DTYPE_BF16 / PCRL_DTYPE_F32) is the storage type of activations and packed
 *     operands: bf16 runs kind::f16 MMAs, fp32 runs kind::tf32 MMAs (fp32 accumulation in both);
 *   - activations are "H-padded NDHWC": logical (N,C,D,H,W) stored as [N][D][H+1][W][C]
 *     with row h'=0 of every plane all zero (voxel h lives at row h+1).  Producers in this
 *     library write that zero row themselves; a caller-made tensor must honour it;
 *   - 1-channel tensors (network input, masks) are plain fp32 [N][D][H][W].
 */
#ifndef PCRL_B200_H
#define PCRL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PCRL_OK 0
#define PCRL_ERR_ARG (-1)
#define PCRL_ERR_CUDA (-2)
#define PCRL_ERR_UNSUPPORTED (-3)

/* storage type of activations and tensor-core operands: bf16 (kind::f16 MMA) or fp32 (kind::tf32) */
#define PCRL_DTYPE_BF16 0
#define PCRL_DTYPE_F32 1
/* fp32 storage WITHOUT the tf32 rounding on store (precision='fp32x3'): activations keep all 24
 * mantissa bits; the tensor-core kernels are then fed operands split by pcrl_split3_tf32 (3xTF32:
 * x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32-equivalent products).  Accepted wherever `dtype` is. */
#define PCRL_DTYPE_F32X 2

/* activation codes: models/pcrlv2_model_3d.py:20-27 */
#define PCRL_ACT_RELU 0
#define PCRL_ACT_PRELU 1
#define PCRL_ACT_ELU 2
#define PCRL_ACT_SIGMOID 3
#define PCRL_ACT_NONE 4
/* nn.LeakyReLU(0.01): not offered by the reference's LUConv (:20-27); BASELINE.json:north_star names the
 * Conv3d+InstanceNorm+LeakyReLU block, so it is available as act='leakyrelu' (an extension) */
#define PCRL_ACT_LEAKYRELU 5

const char* pcrl_last_error(void);
int pcrl_version(void);

/* ---- weight layout converters (state_dict layout <-> tensor-core operand layout) ---------- */
/* nn.Conv3d.weight (Cout,Cin,3,3,3) fp32 -> wf [9 (ky,kx)][3 (kz=2,1,0)][Cout][Cin] bf16 (forward
 * operand) and, if wd != NULL, wd [9][3][Cin][Cout] bf16 with mirrored taps (data-gradient
 * operand).
 * models/pcrlv2_model_3d.py:9 */
int pcrl_pack_conv3_weights(const float* w, void* wf, void* wd, int Cout, int Cin, int dtype,
                            void* stream);
/* packed weight gradient [27][Cout][Cin] fp32 -> (Cout,Cin,3,3,3) fp32 */
int pcrl_unpack_conv3_wgrad(const float* gpk, float* g, int Cout, int Cin, void* stream);
/* nn.ConvTranspose3d.weight (Cin,Cout,2,2,2) fp32 -> wf [(tap,Cout)][Cin] bf16 and
 * wd [Cin][(tap,Cout)] bf16.  models/pcrlv2_model_3d.py:52 */
int pcrl_pack_convT_weights(const float* w, void* wf, void* wd, int Cin, int Cout, int dtype,
                            void* stream);
int pcrl_unpack_convT_wgrad(const float* gpk, float* g, int Cin, int Cout, void* stream);

/* ---- 3x3x3 convolution on tensor cores (tcgen05 implicit GEMM) ---------------------------- */
/* y = conv3d(x, w, padding=1) WITHOUT bias (the bias cancels in the following normalisation and
 * is folded into running_mean by pcrl_norm_finalize).  x [N][D][H+1][W][Cin], y
 * [N][D][H+1][W][Cout] bf16 (out_fp32: fp32).  stats (nullable) [G][Cout][2] fp64 is
 * ACCUMULATED with per-channel sum / sum of squares of the stored y (G = N if stats_per_sample
 * else 1).  Replaces F.conv3d inside LUConv.forward, models/pcrlv2_model_3d.py:33. */
int pcrl_conv3d_k3_fprop(const void* x, const void* wf, void* y, double* stats,
                         int stats_per_sample, int out_fp32, int N, int D, int H, int W, int Cin,
                         int Cout, int dtype, void* stream);
/* dx = conv3d data gradient; dy [N][D][H+1][W][Cout] (pad rows zero), wd from
 * pcrl_pack_conv3_weights, dx [N][D][H+1][W][Cin] bf16.  Autograd of the call above. */
int pcrl_conv3d_k3_dgrad(const void* dy, const void* wd, void* dx, int N, int D, int H, int W,
                         int Cin, int Cout, int dtype, void* stream);
/* Same data gradient, stored for the ConvTranspose3d(k2,s2) that produced x: coarse-major
 * [N*(D/2)*(H/2+1)*(W/2)][8][Cin] bf16 (pad rows zeroed) -- the operand layout of
 * pcrl_convT3d_k2s2_bwd with g_fine = NULL.  colsum (nullable) [Cin][2] fp64 += per-channel
 * (sum, sum of squares) of dx: column 0 is the ConvTranspose bias gradient. */
int pcrl_conv3d_k3_dgrad_unshuffled(const void* dy, const void* wd, void* dx_coarse_major,
                                    double* colsum, int N, int D, int H, int W, int Cin, int Cout,
                                    int dtype, void* stream);
/* dw_packed [27][Cout][Cin] fp32 += weight gradient (caller zeroes or keeps a running sum).
 * dy and x must have zero pad rows. */
int pcrl_conv3d_k3_wgrad(const void* dy, const void* x, float* dw_packed, int N, int D, int H,
                         int W, int Cin, int Cout, int dtype, void* stream);

/* ---- Conv3d(1 -> 32) stem (down_tr64.ops.0), models/pcrlv2_model_3d.py:114 ----------------- */
int pcrl_stem_conv_fprop(const float* x, const float* w, void* y, double* stats,
                         int stats_per_sample, int N, int D, int H, int W, int dtype, void* stream);
/* dw (32,1,3,3,3) fp32 += */
int pcrl_stem_conv_wgrad(const void* dy, const float* x, float* dw, int N, int D, int H, int W,
                         int dtype, void* stream);

/* ---- ConvTranspose3d(k=2, s=2), models/pcrlv2_model_3d.py:52,64 ---------------------------- */
/* y_fine [N][2D][2H+1][2W][Cout] bf16 = convT(x [N][D][H+1][W][Cin]) + bias (pad rows zeroed). */
int pcrl_convT3d_k2s2_fprop(const void* x, const void* wf, const float* bias, void* y_fine, int N,
                            int D, int H, int W, int Cin, int Cout, int dtype, void* stream);
/* Backward.  g_fine [N][2D][2H+1][2W][Cout] bf16; scratch [N*D*(H+1)*W][8*Cout] bf16;
 * dx [N][D][H+1][W][Cin] bf16; dw_packed [(tap,Cout)][Cin] fp32 += ; dbias [Cout] fp32 += .
 * x may be NULL together with dw_packed to skip the weight gradient; g_fine may be NULL when
 * scratch already holds the coarse-major gradient (pcrl_conv3d_k3_dgrad_unshuffled). */
int pcrl_convT3d_k2s2_bwd(const void* g_fine, const void* x, const void* wd, void* scratch,
                          void* dx, float* dw_packed, float* dbias, int N, int D, int H, int W,
                          int Cin, int Cout, int dtype, void* stream);

/* ---- normalisation + activation (+ max-pool, + average-pool sums) -------------------------- */
/* BatchNorm3d / InstanceNorm3d statistics -> scale/shift; updates running stats (BatchNorm) and
 * num_batches_tracked.  models/pcrlv2_model_3d.py:11-16 (F.batch_norm / F.instance_norm). */
int pcrl_norm_finalize(const double* stats, double count, const float* gamma, const float* beta,
                       const float* conv_bias, float* running_mean, float* running_var,
                       long long* num_batches_tracked, float momentum, float eps, float* scale,
                       float* shift, float* mean, float* invstd, int G, int C, void* stream);
/* a = act(y*scale + shift); optional 2x2x2 max-pool (nn.MaxPool3d(2), :100) and per-(n,c) sums
 * for F.adaptive_avg_pool3d (:67). */
int pcrl_norm_act_fwd(const void* y, const float* scale, const float* shift, const float* prelu,
                      void* a_out, void* pool_out, float* avg_sum, int per_sample, int act,
                      int pool, int N, int D, int H, int W, int C, int dtype, void* stream);
/* two-pass backward: pass 0 accumulates sums [G][C][3] fp64, pass 1 writes dy. */
int pcrl_norm_act_bwd(const void* y, const void* g1, const void* g2, const float* gavg,
                      const float* scale, const float* shift, const float* mean,
                      const float* invstd, const float* gamma, const float* prelu, double* sums,
                      void* dy, double count, int per_sample, int act, int pool, int pass, int N,
                      int D, int H, int W, int C, int dtype, void* stream);
/* zero row h'=0 of `planes` planes; row_bytes = bytes of one (w, c) row */
int pcrl_zero_pad_rows(void* t, long long planes, int H1, long long row_bytes, void* stream);

/* ---- single-channel heads ------------------------------------------------------------------ */
/* The N=1 convolutions Conv3d(C->1,k3,p1) (deep_supervision_head.conv1, :60) and Conv3d(64->1,k1)
 * (out_tr.final_conv, :78) are factored as  T = A * Wext^T  (pcrl_gemm_nt, 32 columns: 27 taps,
 * the 1x1x1 conv, 4 zeros) followed by a 27-point gather; backward is the mirror image.
 * w3 (1,C,3,3,3), w1 (1,C,1,1,1) or NULL -> wext [32][C] bf16, wextT [C][32] bf16. */
int pcrl_head_pack_weights(const float* w3, const float* w1, void* wext, void* wextT, int C,
                           int dtype, void* stream);
/* tT [32][rows] fp32 (rows = N*D*(H+1)*W) -> y1 (+ y0) [N][D][H][W] fp32; stats [G][2] fp64 +=
 * (sum, sum of squares) of y1 for the 1-channel norm that follows. */
int pcrl_head_gather(const float* tT, const float* b3, const float* b1, float* y1, float* y0,
                     double* stats, int stats_per_sample, int N, int D, int H, int W, void* stream);
/* dT [rows][32] bf16: dT[u][tap] = dy1[u - tap], dT[u][27] = dy0[u] (dy0 may be NULL). */
int pcrl_head_scatter(const float* dy1, const float* dy0, void* dT, int N, int D, int H, int W,
                      int dtype, void* stream);
/* BatchNorm3d(1)/InstanceNorm3d(1) + Sigmoid of the head (:12,27) on fp32 [G][vol]. */
int pcrl_chan1_sigmoid_fwd(const float* y, const float* scale, const float* shift, float* mask,
                           int per_sample, int G, long long vol, void* stream);
int pcrl_chan1_sigmoid_bwd(const float* y, const float* mask, const float* dmask, const float* mean,
                           const float* invstd, const float* gamma, double* sums, float* dy,
                           double count, int per_sample, int pass, int G, long long vol,
                           void* stream);
/* x [N][D][H][W] fp32 -> X27 [rows][32] bf16, X27[u][tap] = x[u + tap]: im2col of the 1-channel
 * network input; the stem weight gradient is then pcrl_gemm_tn(dY, X27). */
int pcrl_im2col27(const float* x, void* out, int N, int D, int H, int W, int dtype, void* stream);

/* ---- plain tensor-core GEMMs (bf16 or fp32/tf32 in, fp32 accumulate) ------------------------------------ */
/* C[rows][cols] = A[rows][K] * B[cols][K]^T (+ bias[col]); ldc in elements.  out_fp32: 0 = bf16,
 * 1 = fp32, 2 = fp32 transposed (C^T[cols][rows], ldc = rows). */
int pcrl_gemm_nt(const void* a, const void* b, void* c, const float* bias, long long rows, int K,
                 int cols, int ldc, int out_fp32, int dtype, void* stream);
/* C[rows][cols] = A * B^T stored in the operand type, plus stats [cols][2] fp64 += per-column (sum, sum of
 * squares) of the stored C.  With A = pcrl_im2col27(x) and B = the (32, 27 -> 32) stem filter this is the
 * Conv3d(1 -> 32) stem (models/pcrlv2_model_3d.py:114) with the statistics of its BatchNorm, on the tensor
 * cores: the SIMT stem kernel runs at 15 % of the HBM rate, this pair of launches at the copy rate. */
int pcrl_gemm_nt_stats(const void* a, const void* b, void* c, double* stats, long long rows, int K,
                       int cols, int dtype, void* stream);
/* C[P][Q] (fp32) += A[rows][P]^T * B[rows][Q] */
int pcrl_gemm_tn(const void* a, const void* b, float* c, long long rows, int P, int Q, int dtype,
                 void* stream);

/* ---- projection / prediction heads and loss terms (small fp32 tensors) --------------------- */
/* nn.BatchNorm1d over the rows of x [B][C] (UpTransition.bn, predictor_head[1];
 * models/pcrlv2_model_3d.py:54,56,67,69), optionally followed by ReLU (predictor_head[2], :57).
 * training: batch statistics (biased variance), running_mean / running_var / num_batches_tracked
 * updated as torch does (momentum, unbiased variance); eval: running statistics.
 * save_mean / save_invstd [C] are written for the backward pass. */
int pcrl_bn1d_fwd(const float* x, const float* gamma, const float* beta, float* running_mean,
                  float* running_var, long long* num_batches_tracked, float* y, float* save_mean,
                  float* save_invstd, int B, int C, int relu, int training, float momentum, float eps,
                  void* stream);
/* autograd of the above: dx [B][C], dgamma [C], dbeta [C]; y = forward output (ReLU mask) */
int pcrl_bn1d_bwd(const float* x, const float* y, const float* dy, const float* gamma,
                  const float* save_mean, const float* save_invstd, float* dx, float* dgamma,
                  float* dbeta, int B, int C, int relu, int training, void* stream);
/* nn.Linear (predictor_head[0], [3]; :55,58,69): y [B][J] = x [B][K] * w [J][K]^T + bias [J] */
int pcrl_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int K, int J,
                    void* stream);
/* its autograd: dx [B][K] = dy * w, dw [J][K] = dy^T * x, dbias [J] = column sums (each may be NULL) */
int pcrl_linear_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw,
                    float* dbias, int B, int K, int J, void* stream);
/* nn.CosineSimilarity(dim=1)(x, y).mean() of train_3d.py:90-91 with y detached, forward and
 * gradient in one launch: *mean_out += coef * mean_b cos(x_b, y_b);
 * dx [B][C] = coef * d(mean cos)/dx (dx may be NULL). */
int pcrl_cosine_mean_fwd_bwd(const float* x, const float* y, float* mean_out, float* dx, int B, int C,
                             float eps, float coef, void* stream);
/* nn.MSELoss() of train_3d.py:135,137: *out += mean((p - t)^2) over n elements */
int pcrl_mse_fwd(const float* p, const float* t, float* out, long long n, void* stream);
/* dp = g[0] * 2 (p - t) / n, g = upstream gradient (device scalar) */
int pcrl_mse_bwd(const float* p, const float* t, const float* g, float* dp, long long n, void* stream);
/* the same MSE terms multiplied by a DEVICE scalar `weight` (beta * [drawn scale == this scale] of the
 * deep-supervision term, train_3d.py:136-137): lets a captured CUDA graph of the step evaluate all
 * three scales with the drawn one selected by data */
int pcrl_mse_scaled_fwd(const float* p, const float* t, const float* weight, float* out, long long n,
                        void* stream);
int pcrl_mse_scaled_bwd(const float* p, const float* t, const float* g, const float* weight, float* dp,
                        long long n, void* stream);
/* All 1 + 2*n_local cos_loss terms of one step (train_3d.py:86-92,119,124-133) and their gradients
 * wrt the prediction-head outputs in one launch; the drawn scale of every term is read from device
 * memory (draws[0]: global term; draws[1+2i], draws[2+2i]: decoder 1 / decoder 2 against local view
 * i).  ptrs: HOST array of 27 device pointers, scale-major triples in the order pre1, pro1, pre2,
 * pro2 ([B][C_s]), preL, proL ([n_local*B][C_s]), dpre1, dpre2, dpreL (outputs, fully written);
 * channels: HOST int[3] (C_s <= 256).  out2[0] += loss2, out2[1] += local_loss. */
int pcrl_contrastive_fwd_bwd(const void* const* ptrs, const int* channels, int B, int n_local,
                             const int* draws, float* out2, float eps, void* stream);
/* the same for S scales (1..5; the 2-D model has five decoder blocks, train_2d.py:111-117,143-155): ptrs = 9*S
 * pointers in the same group order, channels HOST int[S] */
int pcrl_contrastive_fwd_bwd_s(const void* const* ptrs, const int* channels, int S, int B, int n_local,
                               const int* draws, float* out2, float eps, void* stream);
/* torch.sigmoid of the 1-channel output volume (models/pcrlv2_model_3d.py:79,132) and its autograd */
int pcrl_sigmoid_fwd(const float* x, float* y, long long n, void* stream);
int pcrl_sigmoid_bwd(const float* y, const float* dy, float* dx, long long n, void* stream);
/* F.interpolate(scale_factor=sf, mode='trilinear') of a 1-channel volume x [N][D][H][W] -> y
 * [N][D*sf][H*sf][W*sf] (deep-supervision masks, models/pcrlv2_model_3d.py:125-126); the backward
 * scatters dy into dx, which must be zero on entry */
int pcrl_upsample_trilinear_fwd(const float* x, float* y, int N, int D, int H, int W, int sf,
                                void* stream);
int pcrl_upsample_trilinear_bwd(const float* dy, float* dx, int N, int D, int H, int W, int sf,
                                void* stream);

/* ---- 3xTF32 operand split (precision='fp32x3': fp32-equivalent tensor-core products) ----------- */
/* src [rows][C] fp32 -> three tf32-representable parts, hi = rna_tf32(x), lo = rna_tf32(x - hi):
 *   pattern 0 (the A / activation side):  (hi, lo, hi)     pattern 1 (the B / weight side): (hi, hi, lo)
 *   stack_rows = 0: dst [rows][3*C], the parts concatenated along the contraction (K) index, for
 *                   the K-major kernels (conv fprop / dgrad, gemm_nt): sum_k A3*B3 = hi*hi + lo*hi + hi*lo;
 *   stack_rows = 1: dst [3][rows][C], the parts stacked along the row index, for the kernels that
 *                   reduce over rows (conv wgrad over N*voxels, gemm_tn).
 * The reference computes these products in fp32 (models/pcrlv2_model_3d.py:9,52,60,78 under
 * allow_tf32=False); this is how the same precision is reached on tf32 tensor cores. */
int pcrl_split3_tf32(const float* src, float* dst, long long rows, int C, int pattern, int stack_rows,
                     void* stream);

/* ---- optimizer: torch.optim.SGD(momentum, weight_decay), train_3d.py:48-51,151 ------------- */
int pcrl_sgd_flat(float* params, const float* grads, float* momentum_buf,
                  const long long* seg_offsets, const int* seg_active, const int* seg_first,
                  int nseg, float lr, float momentum, float weight_decay, float grad_scale,
                  void* stream);

/* ---- GPU-side input staging: the per-item torchio transforms of data.py:73-89 /
 * datasets/lunaDataset.py:28-81 as batched kernels over [B][D][H][W] fp32 volumes.  Random parameters are
 * drawn by the host and passed in (device arrays of B entries). ------------------------------------- */
/* torchio RandomFlip: bit 0 / 1 / 2 of axis_mask[b] mirrors axis D / H / W */
int pcrl_aug_flip(const float* x, float* y, const int* axis_mask, int B, int D, int H, int W, void* stream);
/* torchio RandomBlur = scipy.ndimage.gaussian_filter per axis: radius int(4 sigma + 0.5), fp64 weights and
 * accumulation, boundary 'reflect'; sigma[b*sigma_stride + axis]; axis 0 / 1 / 2 = D / H / W; out of place */
int pcrl_aug_blur_axis(const float* x, float* y, const float* sigma, int sigma_stride, int axis, int B,
                       int D, int H, int W, void* stream);
/* torchio RandomNoise then RandomGamma: t = x + noise_std[b] * n, y = sign(t) |t|^exp(log_gamma[b]); n from
 * `noise` [B][vol] when given, else from a counter-based generator keyed by (seed, b, voxel) */
int pcrl_aug_noise_gamma(const float* x, float* y, const float* noise, const float* noise_std,
                         const float* log_gamma, unsigned long long seed, int B, int vol, void* stream);
/* torchio RandomSwap(patch_size, num_iterations): corners [B][iters][6] = first / second patch origin (d,h,w);
 * swaps are applied in order, in place */
int pcrl_aug_swap(float* x, const int* corners, int iters, int pd, int ph, int pw, int B, int D, int H,
                  int W, void* stream);
/* torchio ZNormalization: (x - mean) / std per volume (Bessel-corrected std) */
int pcrl_aug_znorm(const float* x, float* y, int B, int vol, void* stream);

/* ---- offline crop generator pieces ---------------------------------------------- */
/* HU window of luna_preprocess.py:133-135: y = (clip(x, hu_min, hu_max) - hu_min) / (hu_max - hu_min), fp64 inside */
int pcrl_hu_window(const float* x, float* y, long long n, double hu_min, double hu_max, void* stream);
/* the voxel loops of luna_preprocess.py:217-236: crop [X][Y][z_pitch] (z fastest, z_pitch >= Z + len_depth - 1);
 * per voxel (i, j, d < Z) the first k < len_depth with crop[i][j][d+k] >= threshold gives t_img = that value and
 * d_img = 1 - k / (len_depth - 1) (none: t_img = 0, d_img = 0); *sum += sum(d_img) (the lung test of :243-247) */
int pcrl_depth_scan(const float* crop, float* t_img, float* d_img, double* sum, int X, int Y, int Z,
                    int z_pitch, int len_depth, float threshold, void* stream);

/* the same update with the scalars in DEVICE memory, hyper = [lr, momentum, weight_decay, grad_scale,
 * skip_threshold] (utils.py:111-114 changes lr per epoch; a captured graph must not bake it in).
 * guard (nullable): device scalar = loss summed over the ranks; nothing is updated when
 * guard[0]*grad_scale > skip_threshold (train_3d.py:140-142, same decision on every rank). */
int pcrl_sgd_flat_dev(float* params, const float* grads, float* momentum_buf,
                      const long long* seg_offsets, const int* seg_active, const int* seg_first,
                      int nseg, const float* hyper, const float* guard, void* stream);

/* ---- 2-D path: reference models/pcrlv2_model.py + the torchvision ResNet-18 its smp encoder
 * wraps.  Activations are H-padded NHWC = the layout above with D = 1: [N][H+1][W][C].  Convolutions run as
 * im2col -> pcrl_gemm_nt(_stats) / pcrl_gemm_tn -> col2im; `dtype` is PCRL_DTYPE_BF16 or PCRL_DTYPE_F32. ---- */
/* col[(n,ho',wo)][(ky*k+kx)*C + c] = x[n][ho*stride-pad+ky][wo*stride-pad+kx][c] (zero outside the image, in row
 * ho' = 0 and in the K padding kk*C..Kp-1).  image_nchw: x is the network input, fp32 [N][C][H][W] (any C);
 * otherwise x is an H-padded NHWC activation with C % 8 == 0.  F.conv2d of torchvision resnet.py (conv1 7x7/2,
 * the 3x3/2 and 1x1/2 convolutions) and of md.Conv2dReLU, models/pcrlv2_model.py:51-64,78-93 */
int pcrl_im2col2d(const void* x, void* col, int N, int H, int W, int C, int k, int stride, int pad, int Ho,
                  int Wo, int Kp, int image_nchw, int dtype, void* stream);
/* data gradient of the same convolution from dcol = dY * W: dx [N][H+1][W][C] (gather form, no atomics) */
int pcrl_col2im2d(const void* dcol, void* dx, int N, int H, int W, int C, int k, int stride, int pad, int Ho,
                  int Wo, int Kp, int dtype, void* stream);
/* nn.MaxPool2d(3, stride 2, padding 1) of the ResNet stem and its backward (first maximum wins, as torch);
 * y / dy [N][(H-1)/2+2][(W-1)/2+1][C] */
int pcrl_maxpool2d_3x3s2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream);
int pcrl_maxpool2d_3x3s2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int dtype,
                             void* stream);
/* residual join of BasicBlock (`out += identity; relu`): op 0: out = relu(a+b); op 1: out = b * (a > 0)
 * (a = forward output, b = incoming gradient); op 2: out = a + b.  n elements, n % 8 == 0 */
int pcrl_add_relu(const void* a, const void* b, void* out, long long n, int op, int dtype, void* stream);
/* F.interpolate(scale_factor=2, mode='nearest'), models/pcrlv2_model.py:114: x [N][H+1][W][C] ->
 * y [N][2H+1][2W][C]; backward: g fine -> dx coarse (sum of the four children) */
int pcrl_upsample_nearest2x_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream);
int pcrl_upsample_nearest2x_bwd(const void* g, void* dx, int N, int H, int W, int C, int dtype, void* stream);
/* F.interpolate(scale_factor=s, mode='bilinear') (align_corners=False) of the deep-supervision masks,
 * models/pcrlv2_model.py:192: fp32 [NC][H][W] -> [NC][H*s][W*s]; backward ADDS into dx (zeroed by the caller) */
int pcrl_bilinear2d_fwd(const float* x, float* y, int NC, int H, int W, int scale, void* stream);
int pcrl_bilinear2d_bwd(const float* g, float* dx, int NC, int H, int W, int scale, void* stream);
/* Conv2d(C -> 3, k = 1 or 3, padding k/2) with bias: the deep-supervision output conv (models/pcrlv2_model.py:106)
 * and smp's segmentation head (:208).  a [N][H+1][W][Cs] (channels 0..C-1 used), w fp32 [3][C][k][k] (state_dict
 * layout), out fp32 NCHW [N][3][H][W].  Backward: da (nullable) [N][H+1][W][Cs], dw [3][C][k][k] and db [3]
 * (nullable together; ADDED to, zeroed by the caller) */
int pcrl_conv2d_c3_fwd(const void* a, const float* w, const float* bias, float* out, int N, int H, int W, int C,
                       int Cs, int k, int dtype, void* stream);
int pcrl_conv2d_c3_bwd(const void* a, const float* w, const float* dout, void* da, float* dw, float* db, int N,
                       int H, int W, int C, int Cs, int k, int dtype, void* stream);

/* nn.Conv2d.weight (Cout,Cin,k,k) fp32 -> the GEMM operands of the im2col convolution: wmat [CoutP][Kp] (forward)
 * and wt [Kp][CoutP] (data gradient), K index = (ky*k+kx)*cs + c, zero padding, rounded to the operand type;
 * and the way back for the weight gradient the GEMM produced ([CoutP][Kp] fp32, or its transpose) */
int pcrl_pack_conv2d_weights(const float* w, void* wmat, void* wt, int Cout, int Cin, int k, int cs, int CoutP,
                             int Kp, int dtype, void* stream);
int pcrl_unpack_conv2d_wgrad(const float* dw_gemm, float* g, int Cout, int Cin, int k, int cs, int CoutP, int Kp,
                             int transposed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCRL_B200_H */
