// extern "C" surface of libpcrl_b200.so (see include/pcrl_b200.h).  Thin: argument checks and
// forwarding to the kernel launchers in the other translation units.
#include "common.cuh"
#include "../../include/pcrl_b200.h"

namespace pcrl {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// launchers defined in igemm_kmajor.cu / igemm_mnmajor.cu / streaming.cu / heads.cu
int conv3d_k3_igemm(const void*, const void*, void*, double*, int, int, int, int, int, int, int, int, cudaStream_t, int, int);
int gemm_nt_igemm(const void*, const void*, void*, const float*, long long, int, int, int, int, int, int, int, int, int, cudaStream_t, int, double*);
int conv3d_k3_wgrad_igemm(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t, int);
int gemm_tn_igemm(const void*, const void*, float*, long long, int, int, cudaStream_t, int);
int pack_conv3_weights(const float*, void*, void*, int, int, int, cudaStream_t);
int unpack_conv3_wgrad(const float*, float*, int, int, cudaStream_t);
int pack_convT_weights(const float*, void*, void*, int, int, int, cudaStream_t);
int unpack_convT_wgrad(const float*, float*, int, int, cudaStream_t);
int stem_conv_fprop(const float*, const float*, void*, double*, int, int, int, int, int, int, cudaStream_t);
int stem_conv_wgrad(const void*, const float*, float*, int, int, int, int, int, cudaStream_t);
int norm_finalize(const double*, double, const float*, const float*, const float*, float*, float*, long long*, float, float, float*, float*, float*, float*, int, int, cudaStream_t);
int norm_act_fwd(const void*, const float*, const float*, const float*, void*, void*, float*, int, int, int, int, int, int, int, int, int, cudaStream_t);
int norm_act_bwd(const void*, const void*, const void*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, double*, void*, double, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int zero_pad_rows(void*, long long, int, long long, cudaStream_t);
int convT_unshuffle(const void*, void*, float*, int, int, int, int, int, int, cudaStream_t);
int head_pack_weights(const float*, const float*, void*, void*, int, int, cudaStream_t);
int head_gather(const float*, const float*, const float*, float*, float*, double*, int, int, int, int, int, cudaStream_t);
int head_scatter(const float*, const float*, void*, int, int, int, int, int, cudaStream_t);
int im2col27(const float*, void*, int, int, int, int, int, cudaStream_t);
int chan1_sigmoid_fwd(const float*, const float*, const float*, float*, int, int, long long, cudaStream_t);
int chan1_sigmoid_bwd(const float*, const float*, const float*, const float*, const float*, const float*, double*, float*, double, int, int, int, long long, cudaStream_t);
int sgd_flat(float*, const float*, float*, const long long*, const int*, const int*, int, float, float, float, float, cudaStream_t);
int split3_tf32(const float*, float*, long long, int, int, int, cudaStream_t);
// losses.cu
int bn1d_fwd(const float*, const float*, const float*, float*, float*, long long*, float*, float*, float*, int, int, int, int, float, float, cudaStream_t);
int bn1d_bwd(const float*, const float*, const float*, const float*, const float*, const float*, float*, float*, float*, int, int, int, int, cudaStream_t);
int linear_fwd(const float*, const float*, const float*, float*, int, int, int, cudaStream_t);
int linear_bwd(const float*, const float*, const float*, float*, float*, float*, int, int, int, cudaStream_t);
int cosine_mean_fwd_bwd(const float*, const float*, float*, float*, int, int, float, float, cudaStream_t);
int mse_fwd(const float*, const float*, float*, long long, const float*, cudaStream_t);
int mse_bwd(const float*, const float*, const float*, float*, long long, const float*, cudaStream_t);
int contrastive_fwd_bwd(const void* const*, const int*, int, int, const int*, float*, float, cudaStream_t);
int contrastive_fwd_bwd_s(const void* const*, const int*, int, int, int, const int*, float*, float, cudaStream_t);
int sgd_flat_dev(float*, const float*, float*, const long long*, const int*, const int*, int, const float*, const float*, cudaStream_t);
int sigmoid_fwd(const float*, float*, long long, cudaStream_t);
int sigmoid_bwd(const float*, const float*, float*, long long, cudaStream_t);
int upsample_trilinear_fwd(const float*, float*, int, int, int, int, int, cudaStream_t);
int upsample_trilinear_bwd(const float*, float*, int, int, int, int, int, cudaStream_t);

// augment.cu
int aug_flip(const float*, float*, const int*, int, int, int, int, cudaStream_t);
int aug_blur_axis(const float*, float*, const float*, int, int, int, int, int, int, cudaStream_t);
int aug_noise_gamma(const float*, float*, const float*, const float*, const float*, unsigned long long, int, int, cudaStream_t);
int aug_swap(float*, const int*, int, int, int, int, int, int, int, int, cudaStream_t);
int aug_znorm(const float*, float*, int, int, cudaStream_t);
int hu_window(const float*, float*, long long, double, double, cudaStream_t);
int depth_scan(const float*, float*, float*, double*, int, int, int, int, int, float, cudaStream_t);


// planar.cu (2-D path)
int im2col2d(const void*, void*, int, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int col2im2d(const void*, void*, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int maxpool2d_3x3s2(const void*, const void*, void*, int, int, int, int, int, int, cudaStream_t);
int add_relu(const void*, const void*, void*, long long, int, int, cudaStream_t);
int upsample_nearest2x(const void*, void*, int, int, int, int, int, int, cudaStream_t);
int bilinear2d(const float*, float*, int, int, int, int, int, cudaStream_t);
int conv2d_c3_fwd(const void*, const float*, const float*, float*, int, int, int, int, int, int, int, cudaStream_t);
int conv2d_c3_bwd(const void*, const float*, const float*, void*, float*, float*, int, int, int, int, int, int, int, cudaStream_t);
int pack_conv2d_weights(const float*, void*, void*, int, int, int, int, int, int, int, cudaStream_t);
int unpack_conv2d_wgrad(const float*, float*, int, int, int, int, int, int, int, cudaStream_t);

}  // namespace pcrl

using namespace pcrl;
#define ST(s) ((cudaStream_t)(s))
#define NONNULL(p) PCRL_REQUIRE((p) != nullptr, "%s: argument %s is NULL", __func__, #p)

extern "C" {

const char* pcrl_last_error(void) { return last_error_buf(); }
int pcrl_version(void) { return 102; }
static inline int esz(int dtype) { return dtype == PCRL_DTYPE_BF16 ? 2 : 4; }
#define CHECK_DTYPE(d) PCRL_REQUIRE((d) == PCRL_DTYPE_BF16 || (d) == PCRL_DTYPE_F32 || (d) == PCRL_DTYPE_F32X, "%s: unknown dtype %d", __func__, (d))

int pcrl_pack_conv3_weights(const float* w, void* wf, void* wd, int Cout, int Cin, int dtype, void* stream) {
  NONNULL(w); NONNULL(wf); CHECK_DTYPE(dtype);
  return pack_conv3_weights(w, wf, wd, Cout, Cin, dtype, ST(stream));
}
int pcrl_unpack_conv3_wgrad(const float* gpk, float* g, int Cout, int Cin, void* stream) {
  NONNULL(gpk); NONNULL(g);
  return unpack_conv3_wgrad(gpk, g, Cout, Cin, ST(stream));
}
int pcrl_pack_convT_weights(const float* w, void* wf, void* wd, int Cin, int Cout, int dtype, void* stream) {
  NONNULL(w); NONNULL(wf); NONNULL(wd); CHECK_DTYPE(dtype);
  return pack_convT_weights(w, wf, wd, Cin, Cout, dtype, ST(stream));
}
int pcrl_unpack_convT_wgrad(const float* gpk, float* g, int Cin, int Cout, void* stream) {
  NONNULL(gpk); NONNULL(g);
  return unpack_convT_wgrad(gpk, g, Cin, Cout, ST(stream));
}

int pcrl_conv3d_k3_fprop(const void* x, const void* wf, void* y, double* stats, int stats_per_sample,
                         int out_fp32, int N, int D, int H, int W, int Cin, int Cout, int dtype, void* stream) {
  NONNULL(x); NONNULL(wf); NONNULL(y); CHECK_DTYPE(dtype);
  return conv3d_k3_igemm(x, wf, y, stats, stats_per_sample, out_fp32, N, D, H, W, Cin, Cout, ST(stream), 0, dtype);
}
int pcrl_conv3d_k3_dgrad(const void* dy, const void* wd, void* dx, int N, int D, int H, int W,
                         int Cin, int Cout, int dtype, void* stream) {
  NONNULL(dy); NONNULL(wd); NONNULL(dx); CHECK_DTYPE(dtype);
  // the data gradient is a 3x3x3 convolution of dy (Cout channels) with the mirrored, transposed
  // filter: same kernel with the channel roles swapped
  return conv3d_k3_igemm(dy, wd, dx, nullptr, 0, 0, N, D, H, W, Cout, Cin, ST(stream), 0, dtype);
}
int pcrl_conv3d_k3_dgrad_unshuffled(const void* dy, const void* wd, void* dx_coarse_major, double* colsum,
                                    int N, int D, int H, int W, int Cin, int Cout, int dtype, void* stream) {
  NONNULL(dy); NONNULL(wd); NONNULL(dx_coarse_major); CHECK_DTYPE(dtype);
  int rc = conv3d_k3_igemm(dy, wd, dx_coarse_major, colsum, 0, 0, N, D, H, W, Cout, Cin, ST(stream), 1, dtype);
  if (rc) return rc;
  // pad rows (coarse h' = 0) of the coarse-major tensor feed the ConvTranspose GEMMs: zero them
  return zero_pad_rows(dx_coarse_major, (long long)N * (D / 2), H / 2 + 1,
                       (long long)(W / 2) * 8 * Cin * esz(dtype), ST(stream));
}
int pcrl_conv3d_k3_wgrad(const void* dy, const void* x, float* dw_packed, int N, int D, int H, int W,
                         int Cin, int Cout, int dtype, void* stream) {
  NONNULL(dy); NONNULL(x); NONNULL(dw_packed); CHECK_DTYPE(dtype);
  return conv3d_k3_wgrad_igemm(dy, x, dw_packed, N, D, H, W, Cin, Cout, ST(stream), dtype);
}

int pcrl_stem_conv_fprop(const float* x, const float* w, void* y, double* stats, int stats_per_sample,
                         int N, int D, int H, int W, int dtype, void* stream) {
  NONNULL(x); NONNULL(w); NONNULL(y); CHECK_DTYPE(dtype);
  return stem_conv_fprop(x, w, y, stats, stats_per_sample, N, D, H, W, dtype, ST(stream));
}
int pcrl_stem_conv_wgrad(const void* dy, const float* x, float* dw, int N, int D, int H, int W,
                         int dtype, void* stream) {
  NONNULL(dy); NONNULL(x); NONNULL(dw); CHECK_DTYPE(dtype);
  return stem_conv_wgrad(dy, x, dw, N, D, H, W, dtype, ST(stream));
}

int pcrl_convT3d_k2s2_fprop(const void* x, const void* wf, const float* bias, void* y_fine, int N,
                            int D, int H, int W, int Cin, int Cout, int dtype, void* stream) {
  NONNULL(x); NONNULL(wf); NONNULL(y_fine); CHECK_DTYPE(dtype);
  const long long rows = (long long)N * D * (H + 1) * W;
  int rc = gemm_nt_igemm(x, wf, y_fine, bias, rows, Cin, 8 * Cout, Cout, dtype != PCRL_DTYPE_BF16, /*OUT_CONVT*/ 2,
                         D, H, W, Cout, ST(stream), dtype, nullptr);
  if (rc) return rc;
  return zero_pad_rows(y_fine, (long long)N * 2 * D, 2 * H + 1, (long long)2 * W * Cout * esz(dtype), ST(stream));
}
int pcrl_convT3d_k2s2_bwd(const void* g_fine, const void* x, const void* wd, void* scratch, void* dx,
                          float* dw_packed, float* dbias, int N, int D, int H, int W, int Cin,
                          int Cout, int dtype, void* stream) {
  NONNULL(scratch); CHECK_DTYPE(dtype);
  const long long rows = (long long)N * D * (H + 1) * W;
  int rc = PCRL_OK;
  if (g_fine) {   // otherwise `scratch` already holds the coarse-major gradient
    rc = convT_unshuffle(g_fine, scratch, dbias, N, D, H, W, Cout, dtype, ST(stream));
    if (rc) return rc;
  }
  if (dx) {
    NONNULL(wd);
    rc = gemm_nt_igemm(scratch, wd, dx, nullptr, rows, 8 * Cout, Cin, Cin, dtype != PCRL_DTYPE_BF16, /*OUT_ROWS*/ 1,
                       0, 0, 0, 0, ST(stream), dtype, nullptr);
    if (rc) return rc;
  }
  if (dw_packed) {
    NONNULL(x);
    rc = gemm_tn_igemm(scratch, x, dw_packed, rows, 8 * Cout, Cin, ST(stream), dtype);
    if (rc) return rc;
  }
  return PCRL_OK;
}

int pcrl_norm_finalize(const double* stats, double count, const float* gamma, const float* beta,
                       const float* conv_bias, float* running_mean, float* running_var,
                       long long* nbt, float momentum, float eps, float* scale, float* shift,
                       float* mean, float* invstd, int G, int C, void* stream) {
  NONNULL(stats); NONNULL(gamma); NONNULL(beta); NONNULL(scale); NONNULL(shift); NONNULL(mean); NONNULL(invstd);
  return norm_finalize(stats, count, gamma, beta, conv_bias, running_mean, running_var, nbt, momentum,
                       eps, scale, shift, mean, invstd, G, C, ST(stream));
}
int pcrl_norm_act_fwd(const void* y, const float* scale, const float* shift, const float* prelu,
                      void* a_out, void* pool_out, float* avg_sum, int per_sample, int act, int pool,
                      int N, int D, int H, int W, int C, int dtype, void* stream) {
  NONNULL(y); NONNULL(scale); NONNULL(shift); CHECK_DTYPE(dtype);
  return norm_act_fwd(y, scale, shift, prelu, a_out, pool_out, avg_sum, per_sample, act, pool, N, D, H,
                      W, C, dtype, ST(stream));
}
int pcrl_norm_act_bwd(const void* y, const void* g1, const void* g2, const float* gavg,
                      const float* scale, const float* shift, const float* mean, const float* invstd,
                      const float* gamma, const float* prelu, double* sums, void* dy, double count,
                      int per_sample, int act, int pool, int pass, int N, int D, int H, int W, int C,
                      int dtype, void* stream) {
  NONNULL(y); NONNULL(scale); NONNULL(shift); NONNULL(mean); NONNULL(invstd); NONNULL(sums); CHECK_DTYPE(dtype);
  if (pass == 1) { NONNULL(dy); NONNULL(gamma); }
  return norm_act_bwd(y, g1, g2, gavg, scale, shift, mean, invstd, gamma, prelu, sums, dy, count,
                      per_sample, act, pool, pass, N, D, H, W, C, dtype, ST(stream));
}
int pcrl_zero_pad_rows(void* t, long long planes, int H1, long long row_bytes, void* stream) {
  NONNULL(t);
  return zero_pad_rows(t, planes, H1, row_bytes, ST(stream));
}

int pcrl_head_pack_weights(const float* w3, const float* w1, void* wext, void* wextT, int C, int dtype, void* stream) {
  NONNULL(w3); NONNULL(wext); NONNULL(wextT); CHECK_DTYPE(dtype);
  return head_pack_weights(w3, w1, wext, wextT, C, dtype, ST(stream));
}
int pcrl_head_gather(const float* tT, const float* b3, const float* b1, float* y1, float* y0, double* stats,
                     int stats_per_sample, int N, int D, int H, int W, void* stream) {
  NONNULL(tT); NONNULL(b3); NONNULL(y1);
  return head_gather(tT, b3, b1, y1, y0, stats, stats_per_sample, N, D, H, W, ST(stream));
}
int pcrl_head_scatter(const float* dy1, const float* dy0, void* dT, int N, int D, int H, int W, int dtype, void* stream) {
  NONNULL(dy1); NONNULL(dT); CHECK_DTYPE(dtype);
  return head_scatter(dy1, dy0, dT, N, D, H, W, dtype, ST(stream));
}
int pcrl_chan1_sigmoid_fwd(const float* y, const float* scale, const float* shift, float* mask, int per_sample,
                           int G, long long vol, void* stream) {
  NONNULL(y); NONNULL(scale); NONNULL(shift); NONNULL(mask);
  return chan1_sigmoid_fwd(y, scale, shift, mask, per_sample, G, vol, ST(stream));
}
int pcrl_chan1_sigmoid_bwd(const float* y, const float* mask, const float* dmask, const float* mean,
                           const float* invstd, const float* gamma, double* sums, float* dy, double count,
                           int per_sample, int pass, int G, long long vol, void* stream) {
  NONNULL(y); NONNULL(mask); NONNULL(dmask); NONNULL(mean); NONNULL(invstd); NONNULL(sums);
  if (pass == 1) { NONNULL(dy); NONNULL(gamma); }
  return chan1_sigmoid_bwd(y, mask, dmask, mean, invstd, gamma, sums, dy, count, per_sample, pass, G, vol, ST(stream));
}
int pcrl_im2col27(const float* x, void* out, int N, int D, int H, int W, int dtype, void* stream) {
  NONNULL(x); NONNULL(out); CHECK_DTYPE(dtype);
  return im2col27(x, out, N, D, H, W, dtype, ST(stream));
}

int pcrl_gemm_nt(const void* a, const void* b, void* c, const float* bias, long long rows, int K,
                 int cols, int ldc, int out_fp32, int dtype, void* stream) {
  NONNULL(a); NONNULL(b); NONNULL(c); CHECK_DTYPE(dtype);
  return gemm_nt_igemm(a, b, c, bias, rows, K, cols, ldc, out_fp32 != 0, out_fp32 == 2 ? /*OUT_ROWS_T*/ 3 : /*OUT_ROWS*/ 1,
                       0, 0, 0, 0, ST(stream), dtype, nullptr);
}
int pcrl_gemm_nt_stats(const void* a, const void* b, void* c, double* stats, long long rows, int K, int cols,
                       int dtype, void* stream) {
  NONNULL(a); NONNULL(b); NONNULL(c); NONNULL(stats); CHECK_DTYPE(dtype);
  return gemm_nt_igemm(a, b, c, nullptr, rows, K, cols, cols, dtype != PCRL_DTYPE_BF16, /*OUT_ROWS*/ 1, 0, 0, 0, 0,
                       ST(stream), dtype, stats);
}
int pcrl_gemm_tn(const void* a, const void* b, float* c, long long rows, int P, int Q, int dtype, void* stream) {
  NONNULL(a); NONNULL(b); NONNULL(c); CHECK_DTYPE(dtype);
  return gemm_tn_igemm(a, b, c, rows, P, Q, ST(stream), dtype);
}

int pcrl_bn1d_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                  long long* num_batches_tracked, float* y, float* save_mean, float* save_invstd, int B, int C,
                  int relu, int training, float momentum, float eps, void* stream) {
  NONNULL(x); NONNULL(gamma); NONNULL(beta); NONNULL(y); NONNULL(save_mean); NONNULL(save_invstd);
  return bn1d_fwd(x, gamma, beta, running_mean, running_var, num_batches_tracked, y, save_mean, save_invstd, B, C,
                  relu, training, momentum, eps, ST(stream));
}
int pcrl_bn1d_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* save_mean,
                  const float* save_invstd, float* dx, float* dgamma, float* dbeta, int B, int C, int relu,
                  int training, void* stream) {
  NONNULL(x); NONNULL(dy); NONNULL(gamma); NONNULL(save_mean); NONNULL(save_invstd); NONNULL(dx); NONNULL(dgamma); NONNULL(dbeta);
  return bn1d_bwd(x, y, dy, gamma, save_mean, save_invstd, dx, dgamma, dbeta, B, C, relu, training, ST(stream));
}
int pcrl_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int K, int J, void* stream) {
  NONNULL(x); NONNULL(w); NONNULL(y);
  return linear_fwd(x, w, bias, y, B, K, J, ST(stream));
}
int pcrl_linear_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias, int B,
                    int K, int J, void* stream) {
  NONNULL(x); NONNULL(w); NONNULL(dy);
  return linear_bwd(x, w, dy, dx, dw, dbias, B, K, J, ST(stream));
}
int pcrl_cosine_mean_fwd_bwd(const float* x, const float* y, float* mean_out, float* dx, int B, int C, float eps,
                             float coef, void* stream) {
  NONNULL(x); NONNULL(y); NONNULL(mean_out);
  return cosine_mean_fwd_bwd(x, y, mean_out, dx, B, C, eps, coef, ST(stream));
}
int pcrl_mse_fwd(const float* p, const float* t, float* out, long long n, void* stream) {
  NONNULL(p); NONNULL(t); NONNULL(out);
  return mse_fwd(p, t, out, n, nullptr, ST(stream));
}
int pcrl_mse_scaled_fwd(const float* p, const float* t, const float* weight, float* out, long long n, void* stream) {
  NONNULL(p); NONNULL(t); NONNULL(weight); NONNULL(out);
  return mse_fwd(p, t, out, n, weight, ST(stream));
}
int pcrl_mse_bwd(const float* p, const float* t, const float* g, float* dp, long long n, void* stream) {
  NONNULL(p); NONNULL(t); NONNULL(g); NONNULL(dp);
  return mse_bwd(p, t, g, dp, n, nullptr, ST(stream));
}
int pcrl_mse_scaled_bwd(const float* p, const float* t, const float* g, const float* weight, float* dp, long long n,
                        void* stream) {
  NONNULL(p); NONNULL(t); NONNULL(g); NONNULL(weight); NONNULL(dp);
  return mse_bwd(p, t, g, dp, n, weight, ST(stream));
}
int pcrl_contrastive_fwd_bwd(const void* const* ptrs, const int* channels, int B, int n_local, const int* draws,
                             float* out2, float eps, void* stream) {
  NONNULL(ptrs); NONNULL(channels); NONNULL(draws); NONNULL(out2);
  return contrastive_fwd_bwd(ptrs, channels, B, n_local, draws, out2, eps, ST(stream));
}
int pcrl_contrastive_fwd_bwd_s(const void* const* ptrs, const int* channels, int S, int B, int n_local,
                               const int* draws, float* out2, float eps, void* stream) {
  NONNULL(ptrs); NONNULL(channels); NONNULL(draws); NONNULL(out2);
  return contrastive_fwd_bwd_s(ptrs, channels, S, B, n_local, draws, out2, eps, ST(stream));
}
int pcrl_sigmoid_fwd(const float* x, float* y, long long n, void* stream) {
  NONNULL(x); NONNULL(y);
  return sigmoid_fwd(x, y, n, ST(stream));
}
int pcrl_sigmoid_bwd(const float* y, const float* dy, float* dx, long long n, void* stream) {
  NONNULL(y); NONNULL(dy); NONNULL(dx);
  return sigmoid_bwd(y, dy, dx, n, ST(stream));
}
int pcrl_upsample_trilinear_fwd(const float* x, float* y, int N, int D, int H, int W, int sf, void* stream) {
  NONNULL(x); NONNULL(y);
  return upsample_trilinear_fwd(x, y, N, D, H, W, sf, ST(stream));
}
int pcrl_upsample_trilinear_bwd(const float* dy, float* dx, int N, int D, int H, int W, int sf, void* stream) {
  NONNULL(dy); NONNULL(dx);
  return upsample_trilinear_bwd(dy, dx, N, D, H, W, sf, ST(stream));
}

int pcrl_split3_tf32(const float* src, float* dst, long long rows, int C, int pattern, int stack_rows, void* stream) {
  NONNULL(src); NONNULL(dst);
  return split3_tf32(src, dst, rows, C, pattern, stack_rows, ST(stream));
}

int pcrl_sgd_flat(float* params, const float* grads, float* momentum_buf, const long long* seg_offsets,
                  const int* seg_active, const int* seg_first, int nseg, float lr, float momentum,
                  float weight_decay, float grad_scale, void* stream) {
  NONNULL(params); NONNULL(grads); NONNULL(momentum_buf); NONNULL(seg_offsets); NONNULL(seg_active); NONNULL(seg_first);
  return sgd_flat(params, grads, momentum_buf, seg_offsets, seg_active, seg_first, nseg, lr, momentum,
                  weight_decay, grad_scale, ST(stream));
}

int pcrl_aug_flip(const float* x, float* y, const int* axis_mask, int B, int D, int H, int W, void* stream) {
  NONNULL(x); NONNULL(y); NONNULL(axis_mask);
  return aug_flip(x, y, axis_mask, B, D, H, W, ST(stream));
}
int pcrl_aug_blur_axis(const float* x, float* y, const float* sigma, int sigma_stride, int axis, int B, int D,
                       int H, int W, void* stream) {
  NONNULL(x); NONNULL(y); NONNULL(sigma);
  return aug_blur_axis(x, y, sigma, sigma_stride, axis, B, D, H, W, ST(stream));
}
int pcrl_aug_noise_gamma(const float* x, float* y, const float* noise, const float* noise_std,
                         const float* log_gamma, unsigned long long seed, int B, int vol, void* stream) {
  NONNULL(x); NONNULL(y); NONNULL(noise_std); NONNULL(log_gamma);
  return aug_noise_gamma(x, y, noise, noise_std, log_gamma, seed, B, vol, ST(stream));
}
int pcrl_aug_swap(float* x, const int* corners, int iters, int pd, int ph, int pw, int B, int D, int H, int W,
                  void* stream) {
  NONNULL(x); NONNULL(corners);
  return aug_swap(x, corners, iters, pd, ph, pw, B, D, H, W, ST(stream));
}
int pcrl_aug_znorm(const float* x, float* y, int B, int vol, void* stream) {
  NONNULL(x); NONNULL(y);
  return aug_znorm(x, y, B, vol, ST(stream));
}

int pcrl_hu_window(const float* x, float* y, long long n, double hu_min, double hu_max, void* stream) {
  NONNULL(x); NONNULL(y);
  return hu_window(x, y, n, hu_min, hu_max, ST(stream));
}
int pcrl_depth_scan(const float* crop, float* t_img, float* d_img, double* sum, int X, int Y, int Z, int z_pitch,
                    int len_depth, float threshold, void* stream) {
  NONNULL(crop); NONNULL(t_img); NONNULL(d_img); NONNULL(sum);
  return depth_scan(crop, t_img, d_img, sum, X, Y, Z, z_pitch, len_depth, threshold, ST(stream));
}

int pcrl_sgd_flat_dev(float* params, const float* grads, float* momentum_buf, const long long* seg_offsets,
                      const int* seg_active, const int* seg_first, int nseg, const float* hyper,
                      const float* guard, void* stream) {
  NONNULL(params); NONNULL(grads); NONNULL(momentum_buf); NONNULL(seg_offsets); NONNULL(seg_active); NONNULL(seg_first);
  NONNULL(hyper);
  return sgd_flat_dev(params, grads, momentum_buf, seg_offsets, seg_active, seg_first, nseg, hyper, guard, ST(stream));
}

// ---- 2-D path (planar.cu)
#define CHECK_DTYPE2(d) CHECK_DTYPE(d)
int pcrl_im2col2d(const void* x, void* col, int N, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                  int Kp, int image_nchw, int dtype, void* stream) {
  NONNULL(x); NONNULL(col); CHECK_DTYPE2(dtype);
  PCRL_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0, "pcrl_im2col2d: bad dims");
  return im2col2d(x, col, N, H, W, C, k, stride, pad, Ho, Wo, Kp, image_nchw, dtype, ST(stream));
}
int pcrl_col2im2d(const void* dcol, void* dx, int N, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                  int Kp, int dtype, void* stream) {
  NONNULL(dcol); NONNULL(dx); CHECK_DTYPE2(dtype);
  PCRL_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0, "pcrl_col2im2d: bad dims");
  return col2im2d(dcol, dx, N, H, W, C, k, stride, pad, Ho, Wo, Kp, dtype, ST(stream));
}
int pcrl_maxpool2d_3x3s2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream) {
  NONNULL(x); NONNULL(y); CHECK_DTYPE2(dtype);
  return maxpool2d_3x3s2(x, nullptr, y, N, H, W, C, 0, dtype, ST(stream));
}
int pcrl_maxpool2d_3x3s2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int dtype,
                             void* stream) {
  NONNULL(x); NONNULL(dy); NONNULL(dx); CHECK_DTYPE2(dtype);
  return maxpool2d_3x3s2(x, dy, dx, N, H, W, C, 1, dtype, ST(stream));
}
int pcrl_add_relu(const void* a, const void* b, void* out, long long n, int op, int dtype, void* stream) {
  NONNULL(a); NONNULL(b); NONNULL(out); CHECK_DTYPE2(dtype);
  return add_relu(a, b, out, n, op, dtype, ST(stream));
}
int pcrl_upsample_nearest2x_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream) {
  NONNULL(x); NONNULL(y); CHECK_DTYPE2(dtype);
  return upsample_nearest2x(x, y, N, H, W, C, 0, dtype, ST(stream));
}
int pcrl_upsample_nearest2x_bwd(const void* g, void* dx, int N, int H, int W, int C, int dtype, void* stream) {
  NONNULL(g); NONNULL(dx); CHECK_DTYPE2(dtype);
  return upsample_nearest2x(g, dx, N, H, W, C, 1, dtype, ST(stream));
}
int pcrl_bilinear2d_fwd(const float* x, float* y, int NC, int H, int W, int scale, void* stream) {
  NONNULL(x); NONNULL(y);
  return bilinear2d(x, y, NC, H, W, scale, 0, ST(stream));
}
int pcrl_bilinear2d_bwd(const float* g, float* dx, int NC, int H, int W, int scale, void* stream) {
  NONNULL(g); NONNULL(dx);
  return bilinear2d(g, dx, NC, H, W, scale, 1, ST(stream));
}
int pcrl_conv2d_c3_fwd(const void* a, const float* w, const float* bias, float* out, int N, int H, int W, int C,
                       int Cs, int k, int dtype, void* stream) {
  NONNULL(a); NONNULL(w); NONNULL(bias); NONNULL(out); CHECK_DTYPE2(dtype);
  return conv2d_c3_fwd(a, w, bias, out, N, H, W, C, Cs, k, dtype, ST(stream));
}
int pcrl_conv2d_c3_bwd(const void* a, const float* w, const float* dout, void* da, float* dw, float* db, int N,
                       int H, int W, int C, int Cs, int k, int dtype, void* stream) {
  NONNULL(a); NONNULL(w); NONNULL(dout); CHECK_DTYPE2(dtype);
  return conv2d_c3_bwd(a, w, dout, da, dw, db, N, H, W, C, Cs, k, dtype, ST(stream));
}

int pcrl_pack_conv2d_weights(const float* w, void* wmat, void* wt, int Cout, int Cin, int k, int cs, int CoutP, int Kp,
                             int dtype, void* stream) {
  NONNULL(w); NONNULL(wmat); NONNULL(wt); CHECK_DTYPE2(dtype);
  return pack_conv2d_weights(w, wmat, wt, Cout, Cin, k, cs, CoutP, Kp, dtype, ST(stream));
}
int pcrl_unpack_conv2d_wgrad(const float* dw_gemm, float* g, int Cout, int Cin, int k, int cs, int CoutP, int Kp,
                             int transposed, void* stream) {
  NONNULL(dw_gemm); NONNULL(g);
  return unpack_conv2d_wgrad(dw_gemm, g, Cout, Cin, k, cs, CoutP, Kp, transposed, ST(stream));
}

}  // extern "C"
