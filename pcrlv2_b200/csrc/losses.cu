// Projection / prediction heads and loss terms of the pre-training step (small fp32 tensors, CUDA
// cores; every reduction is a warp-shuffle tree):
//   * BatchNorm1d (+ReLU) over the rows of a (B, C) matrix        models/pcrlv2_model_3d.py:54,56,67
//   * nn.Linear forward / backward as tiled SGEMMs                 models/pcrlv2_model_3d.py:55,58,68
//   * mean cosine similarity of two (B, C) matrices + gradient     train_3d.py:90-91 (nn.CosineSimilarity)
//   * mean squared error + gradient                                train_3d.py:135,137 (nn.MSELoss)
//   * sigmoid on a 1-channel volume                                models/pcrlv2_model_3d.py:79,132
//   * trilinear x2 / x4 upsampling of a 1-channel volume           models/pcrlv2_model_3d.py:125-126
// B <= a few hundred rows, C <= 512: these kernels are launch-latency sized, not bandwidth sized;
// what matters is that one launch replaces the 5-10 ATen launches of the reference per call.
#include "common.cuh"

namespace pcrl {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// ------------------------------------------------------------------------------ BatchNorm1d
// block = 32 columns x 8 row groups; a column's B values are reduced by its 8 threads through
// shared memory (two passes over the rows: mean, then centred second moment).
template <bool BWD>
__global__ void __launch_bounds__(256)
bn1d_kernel(const float* __restrict__ x, const float* __restrict__ y_fwd, const float* __restrict__ dy,
            const float* __restrict__ gamma, const float* __restrict__ beta,
            float* __restrict__ running_mean, float* __restrict__ running_var,
            long long* __restrict__ nbt, float* __restrict__ out, float* __restrict__ save_mean,
            float* __restrict__ save_invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
            int B, int C, int relu, int training, float momentum, float eps) {
  __shared__ float red[2][8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const bool ok = c < C;
  if (!BWD) {
    float mean, invstd;
    if (training) {
      float s = 0.f;
      if (ok) for (int b = ty; b < B; b += 8) s += x[(size_t)b * C + c];
      red[0][ty][tx] = s;
      __syncthreads();
      s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; i++) s += red[0][i][tx];
      mean = s / (float)B;
      float q = 0.f;
      if (ok) for (int b = ty; b < B; b += 8) { const float d = x[(size_t)b * C + c] - mean; q = fmaf(d, d, q); }
      red[1][ty][tx] = q;
      __syncthreads();
      q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; i++) q += red[1][i][tx];
      const float var = q / (float)B;                    // biased: normalisation
      invstd = 1.f / sqrtf(var + eps);
      if (ok && ty == 0) {
        if (running_mean) {
          running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
          running_var[c] = (1.f - momentum) * running_var[c] + momentum * (q / (float)(B - 1));   // unbiased
        }
        if (nbt && c == 0) *nbt += 1;
      }
    } else {
      mean = ok ? running_mean[c] : 0.f;
      invstd = ok ? 1.f / sqrtf(running_var[c] + eps) : 0.f;
    }
    if (ok) {
      if (ty == 0) { save_mean[c] = mean; save_invstd[c] = invstd; }
      const float g = gamma[c] * invstd, sh = beta[c] - mean * g;
      for (int b = ty; b < B; b += 8) {
        float v = fmaf(x[(size_t)b * C + c], g, sh);
        if (relu) v = fmaxf(v, 0.f);
        out[(size_t)b * C + c] = v;
      }
    }
  } else {
    const float mean = ok ? save_mean[c] : 0.f, invstd = ok ? save_invstd[c] : 0.f;
    float s0 = 0.f, s1 = 0.f;
    if (ok)
      for (int b = ty; b < B; b += 8) {
        const size_t i = (size_t)b * C + c;
        float dz = dy[i];
        if (relu && !(y_fwd[i] > 0.f)) dz = 0.f;
        s0 += dz;
        s1 = fmaf(dz, (x[i] - mean) * invstd, s1);
      }
    red[0][ty][tx] = s0;
    red[1][ty][tx] = s1;
    __syncthreads();
    s0 = s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { s0 += red[0][i][tx]; s1 += red[1][i][tx]; }
    if (ok) {
      if (ty == 0) { dbeta[c] = s0; dgamma[c] = s1; }
      const float g = gamma[c] * invstd;
      const float k0 = training ? s0 / (float)B : 0.f, k1 = training ? s1 / (float)B : 0.f;
      for (int b = ty; b < B; b += 8) {
        const size_t i = (size_t)b * C + c;
        float dz = dy[i];
        if (relu && !(y_fwd[i] > 0.f)) dz = 0.f;
        out[i] = g * (dz - k0 - (x[i] - mean) * invstd * k1);
      }
    }
  }
}

int bn1d_fwd(const float* x, const float* gamma, const float* beta, float* rm, float* rv, long long* nbt,
             float* y, float* save_mean, float* save_invstd, int B, int C, int relu, int training,
             float momentum, float eps, cudaStream_t s) {
  PCRL_REQUIRE(B >= 1 && C >= 1, "bn1d_fwd: bad dims");
  PCRL_REQUIRE(!training || B > 1, "bn1d_fwd: expected more than 1 value per channel when training (B=%d)", B);
  PCRL_REQUIRE(training || (rm && rv), "bn1d_fwd: eval mode needs running statistics");
  bn1d_kernel<false><<<(C + 31) / 32, 256, 0, s>>>(x, nullptr, nullptr, gamma, beta, rm, rv, nbt, y, save_mean,
                                                   save_invstd, nullptr, nullptr, B, C, relu, training, momentum, eps);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int bn1d_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* save_mean,
             const float* save_invstd, float* dx, float* dgamma, float* dbeta, int B, int C, int relu,
             int training, cudaStream_t s) {
  PCRL_REQUIRE(B >= 1 && C >= 1, "bn1d_bwd: bad dims");
  PCRL_REQUIRE(!relu || y, "bn1d_bwd: the ReLU mask needs the forward output");
  bn1d_kernel<true><<<(C + 31) / 32, 256, 0, s>>>(x, y, dy, gamma, nullptr, nullptr, nullptr, nullptr, dx,
                                                  const_cast<float*>(save_mean), const_cast<float*>(save_invstd),
                                                  dgamma, dbeta, B, C, relu, training, 0.f, 0.f);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// ------------------------------------------------------------------------------ small SGEMM
// C[M][N] = op(A) * op(B) (+ bias[n]); A(m,k) = TA ? A[k*lda+m] : A[m*lda+k],
// B(k,n) = TB ? B[n*ldb+k] : B[k*ldb+n].  32x32 tile, 256 threads, 4 outputs per thread.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_tile_kernel(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ bias,
                  float* __restrict__ Cm, int M, int N, int K, int lda, int ldb, int ldc) {
  __shared__ float as[32][33], bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = ty + 8 * i;
      // as[m][k], bs[k][n]; the global index that varies with tx is the contiguous one
      if (TA) { const int k = k0 + r, m = m0 + tx; as[tx][r] = (k < K && m < M) ? A[(size_t)k * lda + m] : 0.f; }
      else    { const int m = m0 + r, k = k0 + tx; as[r][tx] = (k < K && m < M) ? A[(size_t)m * lda + k] : 0.f; }
      if (TB) { const int n = n0 + r, k = k0 + tx; bs[tx][r] = (k < K && n < N) ? Bm[(size_t)n * ldb + k] : 0.f; }
      else    { const int k = k0 + r, n = n0 + tx; bs[r][tx] = (k < K && n < N) ? Bm[(size_t)k * ldb + n] : 0.f; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k++) {
      const float b = bs[k][tx];
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i] = fmaf(as[ty + 8 * i][k], b, acc[i]);
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n < N) {
    const float bv = bias ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int m = m0 + ty + 8 * i;
      if (m < M) Cm[(size_t)m * ldc + n] = acc[i] + bv;
    }
  }
}

// out[n] = sum_m A[m][n]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ A, float* __restrict__ out, int M, int N) {
  __shared__ float red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N) for (int m = ty; m < M; m += 8) s += A[(size_t)m * N + n];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += red[i][tx];
    out[n] = s;
  }
}

int linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int K, int J, cudaStream_t s) {
  PCRL_REQUIRE(B >= 1 && K >= 1 && J >= 1, "linear_fwd: bad dims");
  dim3 grid((J + 31) / 32, (B + 31) / 32);
  sgemm_tile_kernel<false, true><<<grid, 256, 0, s>>>(x, w, bias, y, B, J, K, K, K, J);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int linear_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias,
               int B, int K, int J, cudaStream_t s) {
  PCRL_REQUIRE(B >= 1 && K >= 1 && J >= 1, "linear_bwd: bad dims");
  if (dx) {   // dx[B][K] = dy[B][J] * w[J][K]
    dim3 grid((K + 31) / 32, (B + 31) / 32);
    sgemm_tile_kernel<false, false><<<grid, 256, 0, s>>>(dy, w, nullptr, dx, B, K, J, J, K, K);
    PCRL_CHECK_LAUNCH();
  }
  if (dw) {   // dw[J][K] = dy^T[J][B] * x[B][K]
    dim3 grid((K + 31) / 32, (J + 31) / 32);
    sgemm_tile_kernel<true, false><<<grid, 256, 0, s>>>(dy, x, nullptr, dw, J, K, B, J, K, K);
    PCRL_CHECK_LAUNCH();
  }
  if (dbias) {
    colsum_kernel<<<(J + 31) / 32, 256, 0, s>>>(dy, dbias, B, J);
    PCRL_CHECK_LAUNCH();
  }
  return PCRL_OK;
}

// ------------------------------------------------------------------------------ cosine similarity
// One warp per row.  cos_b = <x_b, y_b> / (max(|x_b|, eps) * max(|y_b|, eps)); *mean_out += coef * mean_b cos_b;
// dx_b = coef/B * (y_b / (|x_b||y_b|) - cos_b * x_b / |x_b|^2)  (y is a constant: the reference detaches it).
__global__ void __launch_bounds__(256)
cosine_mean_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ mean_out,
                   float* __restrict__ dx, int B, int C, float eps, float coef) {
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= B) return;
  const float* xr = x + (size_t)row * C;
  const float* yr = y + (size_t)row * C;
  float dot = 0.f, nx = 0.f, ny = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float a = xr[c], b = yr[c];
    dot = fmaf(a, b, dot);
    nx = fmaf(a, a, nx);
    ny = fmaf(b, b, ny);
  }
  dot = warp_sum(dot); nx = warp_sum(nx); ny = warp_sum(ny);
  const float lx = fmaxf(sqrtf(nx), eps), ly = fmaxf(sqrtf(ny), eps);
  const float inv = 1.f / (lx * ly);
  const float cs = dot * inv;
  const float k = coef / (float)B;
  if (lane == 0) atomicAdd(mean_out, cs * k);
  if (dx) {
    float* dr = dx + (size_t)row * C;
    const float kx = cs / (lx * lx);
    for (int c = lane; c < C; c += 32) dr[c] = k * (yr[c] * inv - xr[c] * kx);
  }
}

int cosine_mean_fwd_bwd(const float* x, const float* y, float* mean_out, float* dx, int B, int C, float eps,
                        float coef, cudaStream_t s) {
  PCRL_REQUIRE(B >= 1 && C >= 1, "cosine_mean: bad dims");
  cosine_mean_kernel<<<(B + 7) / 8, 256, 0, s>>>(x, y, mean_out, dx, B, C, eps, coef);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// ------------------------------------------------------------------------------ fused contrastive terms
// All 1 + 2*n_local cos_loss terms of one training step (train_3d.py:86-92,119,124-133) in ONE launch,
// with the randomly drawn scale of every term read from DEVICE memory (draws[0] = global term, then
// for view i: draws[1+2i] = (decoder 1, view i), draws[2+2i] = (decoder 2, view i)), so that a captured
// CUDA graph of the step is independent of the draws.  One warp per row of a `pre` tensor (the only
// tensors the reference differentiates through: the projections are detached), which loops over the
// terms that involve it: the gradient of a row is written once, without atomics.
//   out[0] += loss2      = -1/2 (mean cos(pre1_s, pro2_s) + mean cos(pre2_s, pro1_s)),  s = draws[0]
//   out[1] += local_loss = 1/(2 n_local) sum_i [ -1/2 (mean cos(pre1_a, proL_a[i]) + mean cos(preL_a[i], pro1_a))
//                                               -1/2 (mean cos(pre2_b, proL_b[i]) + mean cos(preL_b[i], pro2_b)) ]
#define PCRL_MAX_SCALES 5      // 3 (3-D model: up_tr256/128/64) or 5 (2-D model: the five decoder blocks)
struct ContrastiveArgs {
  const float* pre1[PCRL_MAX_SCALES]; const float* pro1[PCRL_MAX_SCALES]; const float* pre2[PCRL_MAX_SCALES];
  const float* pro2[PCRL_MAX_SCALES]; const float* preL[PCRL_MAX_SCALES]; const float* proL[PCRL_MAX_SCALES];
  float* dpre1[PCRL_MAX_SCALES]; float* dpre2[PCRL_MAX_SCALES]; float* dpreL[PCRL_MAX_SCALES];
  int C[PCRL_MAX_SCALES];
  int S;
  int B, n_local;
  const int* draws;
  float* out;
  float eps;
};

__device__ __forceinline__ void cos_term(const float (&xv)[8], float nx, const float* __restrict__ yr, int C, int lane,
                                         float eps, float k, float (&gacc)[8], float& loss) {
  float yv[8], dot = 0.f, ny = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int c = lane + 32 * j;
    yv[j] = c < C ? yr[c] : 0.f;
    dot = fmaf(xv[j], yv[j], dot);
    ny = fmaf(yv[j], yv[j], ny);
  }
  dot = warp_sum(dot); ny = warp_sum(ny);
  const float lx = fmaxf(sqrtf(nx), eps), ly = fmaxf(sqrtf(ny), eps);
  const float inv = 1.f / (lx * ly);
  const float cs = dot * inv;
  const float kx = cs / (lx * lx);
  loss += cs * k;
#pragma unroll
  for (int j = 0; j < 8; j++) gacc[j] = fmaf(k, yv[j] * inv - xv[j] * kx, gacc[j]);
}

__global__ void __launch_bounds__(256)
contrastive_kernel(const ContrastiveArgs a) {
  const int lane = threadIdx.x & 31;
  const int wid = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int B = a.B, nl = a.n_local;
  const int per_scale = B * (2 + nl);
  if (wid >= a.S * per_scale) return;
  const int s = wid / per_scale;
  int r = wid % per_scale;
  const int C = a.C[s];
  const float kg = -0.5f / (float)B, kl = -0.5f / (float)B / (float)(2 * nl);
  const float* xr;
  float* dr;
  int grp, b, view = 0;
  if (r < B) { grp = 0; b = r; xr = a.pre1[s] + (size_t)b * C; dr = a.dpre1[s] + (size_t)b * C; }
  else if (r < 2 * B) { grp = 1; b = r - B; xr = a.pre2[s] + (size_t)b * C; dr = a.dpre2[s] + (size_t)b * C; }
  else { grp = 2; r -= 2 * B; view = r / B; b = r % B; xr = a.preL[s] + (size_t)r * C; dr = a.dpreL[s] + (size_t)r * C; }
  float xv[8], gacc[8], nx = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int c = lane + 32 * j;
    xv[j] = c < C ? xr[c] : 0.f;
    gacc[j] = 0.f;
    nx = fmaf(xv[j], xv[j], nx);
  }
  nx = warp_sum(nx);
  float lg = 0.f, ll = 0.f;
  if (grp == 0) {
    if (a.draws[0] == s) cos_term(xv, nx, a.pro2[s] + (size_t)b * C, C, lane, a.eps, kg, gacc, lg);
    for (int i = 0; i < nl; i++)
      if (a.draws[1 + 2 * i] == s) cos_term(xv, nx, a.proL[s] + ((size_t)i * B + b) * C, C, lane, a.eps, kl, gacc, ll);
  } else if (grp == 1) {
    if (a.draws[0] == s) cos_term(xv, nx, a.pro1[s] + (size_t)b * C, C, lane, a.eps, kg, gacc, lg);
    for (int i = 0; i < nl; i++)
      if (a.draws[2 + 2 * i] == s) cos_term(xv, nx, a.proL[s] + ((size_t)i * B + b) * C, C, lane, a.eps, kl, gacc, ll);
  } else {
    if (a.draws[1 + 2 * view] == s) cos_term(xv, nx, a.pro1[s] + (size_t)b * C, C, lane, a.eps, kl, gacc, ll);
    if (a.draws[2 + 2 * view] == s) cos_term(xv, nx, a.pro2[s] + (size_t)b * C, C, lane, a.eps, kl, gacc, ll);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int c = lane + 32 * j;
    if (c < C) dr[c] = gacc[j];
  }
  if (lane == 0) {
    if (lg != 0.f) atomicAdd(&a.out[0], lg);
    if (ll != 0.f) atomicAdd(&a.out[1], ll);
  }
}

// ptrs: 9*S device pointers in the order pre1[S], pro1[S], pre2[S], pro2[S], preL[S], proL[S], dpre1[S],
// dpre2[S], dpreL[S] (3-D model: scale index 0..2 = up_tr256, up_tr128, up_tr64; 2-D model: decoder blocks 0..4)
int contrastive_fwd_bwd_s(const void* const* ptrs, const int* Cs, int S, int B, int n_local, const int* draws,
                          float* out2, float eps, cudaStream_t s) {
  PCRL_REQUIRE(B >= 1 && n_local >= 1 && S >= 1 && S <= PCRL_MAX_SCALES, "contrastive: bad dims (S=%d)", S);
  ContrastiveArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < S; i++) {
    PCRL_REQUIRE(Cs[i] >= 1 && Cs[i] <= 256, "contrastive: C=%d must be in 1..256", Cs[i]);
    a.C[i] = Cs[i];
    a.pre1[i] = (const float*)ptrs[0 * S + i]; a.pro1[i] = (const float*)ptrs[1 * S + i];
    a.pre2[i] = (const float*)ptrs[2 * S + i]; a.pro2[i] = (const float*)ptrs[3 * S + i];
    a.preL[i] = (const float*)ptrs[4 * S + i]; a.proL[i] = (const float*)ptrs[5 * S + i];
    a.dpre1[i] = (float*)ptrs[6 * S + i]; a.dpre2[i] = (float*)ptrs[7 * S + i]; a.dpreL[i] = (float*)ptrs[8 * S + i];
  }
  for (int i = 0; i < 9 * S; i++) PCRL_REQUIRE(ptrs[i] != nullptr, "contrastive: pointer %d is NULL", i);
  a.S = S; a.B = B; a.n_local = n_local; a.draws = draws; a.out = out2; a.eps = eps;
  const int warps = S * B * (2 + n_local);
  contrastive_kernel<<<(warps + 7) / 8, 256, 0, s>>>(a);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int contrastive_fwd_bwd(const void* const* ptrs, const int* C3, int B, int n_local, const int* draws,
                        float* out2, float eps, cudaStream_t s) {
  return contrastive_fwd_bwd_s(ptrs, C3, 3, B, n_local, draws, out2, eps, s);
}

// ------------------------------------------------------------------------------ MSE
__global__ void __launch_bounds__(256)
mse_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t, float* __restrict__ out, long long n,
               float inv_n, const float* __restrict__ wgt) {
  __shared__ float red[8];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = p4[i], b = t4[i];
    const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
    s = fmaf(d0, d0, s); s = fmaf(d1, d1, s); s = fmaf(d2, d2, s); s = fmaf(d3, d3, s);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float d = p[n4 * 4 + threadIdx.x] - t[n4 * 4 + threadIdx.x];
    s = fmaf(d, d, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int k = 4; k >= 1; k >>= 1) s += __shfl_xor_sync(0xffu, s, k);
    if (threadIdx.x == 0) atomicAdd(out, s * inv_n * (wgt ? wgt[0] : 1.f));
  }
}

__global__ void __launch_bounds__(256)
mse_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ g,
               float* __restrict__ dp, long long n, float two_over_n, const float* __restrict__ wgt) {
  const float k = g[0] * two_over_n * (wgt ? wgt[0] : 1.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dp[i] = k * (p[i] - t[i]);
}

static inline int blocks_for(long long n, int per_block, int cap) {
  long long b = (n + per_block - 1) / per_block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// wgt (nullable): device scalar multiplying the term (beta * one-hot of the drawn deep-supervision scale)
int mse_fwd(const float* p, const float* t, float* out, long long n, const float* wgt, cudaStream_t s) {
  PCRL_REQUIRE(n >= 1, "mse_fwd: empty input");
  PCRL_REQUIRE((((uintptr_t)p | (uintptr_t)t) & 15) == 0, "mse_fwd: inputs must be 16-byte aligned");
  mse_fwd_kernel<<<blocks_for(n, 1024 * 4, num_sms() * 4), 256, 0, s>>>(p, t, out, n, 1.f / (float)n, wgt);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int mse_bwd(const float* p, const float* t, const float* g, float* dp, long long n, const float* wgt, cudaStream_t s) {
  PCRL_REQUIRE(n >= 1, "mse_bwd: empty input");
  mse_bwd_kernel<<<blocks_for(n, 1024, num_sms() * 8), 256, 0, s>>>(p, t, g, dp, n, 2.f / (float)n, wgt);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// ------------------------------------------------------------------------------ sigmoid
__global__ void __launch_bounds__(256)
sigmoid_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n, int bwd) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (!bwd) out[i] = 1.f / (1.f + expf(-a[i]));
    else { const float y = a[i]; out[i] = b[i] * y * (1.f - y); }      // a = y (forward output), b = dy
  }
}

int sigmoid_fwd(const float* x, float* y, long long n, cudaStream_t s) {
  sigmoid_kernel<<<blocks_for(n, 1024, num_sms() * 8), 256, 0, s>>>(x, nullptr, y, n, 0);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int sigmoid_bwd(const float* y, const float* dy, float* dx, long long n, cudaStream_t s) {
  sigmoid_kernel<<<blocks_for(n, 1024, num_sms() * 8), 256, 0, s>>>(y, dy, dx, n, 1);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// ------------------------------------------------------------------------------ trilinear upsample
// F.interpolate(scale_factor=sf, mode='trilinear', align_corners=False) of a 1-channel volume:
// src = max((o + 0.5) / sf - 0.5, 0), i0 = floor(src), i1 = min(i0 + 1, size - 1), lambda = src - i0.
__device__ __forceinline__ void tri_coord(int o, int sf, int size, int& i0, int& i1, float& l1) {
  float src = ((float)o + 0.5f) / (float)sf - 0.5f;
  src = fmaxf(src, 0.f);
  i0 = min((int)src, size - 1);
  i1 = min(i0 + 1, size - 1);
  l1 = src - (float)i0;
}

template <bool BWD>
__global__ void __launch_bounds__(256)
upsample_trilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int D, int H, int W, int sf) {
  const int OD = D * sf, OH = H * sf, OW = W * sf;
  const long long total = (long long)N * OD * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % OW);
    const int oh = (int)((i / OW) % OH);
    const int od = (int)((i / ((long long)OW * OH)) % OD);
    const int n = (int)(i / ((long long)OW * OH * OD));
    int d0, d1, h0, h1, w0, w1;
    float ld, lh, lw;
    tri_coord(od, sf, D, d0, d1, ld);
    tri_coord(oh, sf, H, h0, h1, lh);
    tri_coord(ow, sf, W, w0, w1, lw);
    const size_t base = (size_t)n * D * H * W;
#define IDX(d, h, w) (base + ((size_t)(d) * H + (h)) * W + (w))
    const float wd[2] = {1.f - ld, ld}, wh[2] = {1.f - lh, lh}, ww[2] = {1.f - lw, lw};
    const int dd[2] = {d0, d1}, hh[2] = {h0, h1}, wq[2] = {w0, w1};
    if (!BWD) {
      float v = 0.f;
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
          for (int c = 0; c < 2; c++) v = fmaf(wd[a] * wh[b] * ww[c], src[IDX(dd[a], hh[b], wq[c])], v);
      dst[i] = v;
    } else {
      const float g = src[i];      // src = upstream gradient at the fine voxel, dst = coarse gradient
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
          for (int c = 0; c < 2; c++) atomicAdd(&dst[IDX(dd[a], hh[b], wq[c])], wd[a] * wh[b] * ww[c] * g);
    }
#undef IDX
  }
}

int upsample_trilinear_fwd(const float* x, float* y, int N, int D, int H, int W, int sf, cudaStream_t s) {
  PCRL_REQUIRE(sf >= 1 && N >= 1 && D >= 1 && H >= 1 && W >= 1, "upsample_trilinear_fwd: bad dims");
  const long long total = (long long)N * D * H * W * sf * sf * sf;
  upsample_trilinear_kernel<false><<<blocks_for(total, 512, num_sms() * 16), 256, 0, s>>>(x, y, N, D, H, W, sf);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// dx must be zero on entry (the gradient is scattered with atomics, as ATen's own backward does)
int upsample_trilinear_bwd(const float* dy, float* dx, int N, int D, int H, int W, int sf, cudaStream_t s) {
  PCRL_REQUIRE(sf >= 1 && N >= 1 && D >= 1 && H >= 1 && W >= 1, "upsample_trilinear_bwd: bad dims");
  const long long total = (long long)N * D * H * W * sf * sf * sf;
  upsample_trilinear_kernel<true><<<blocks_for(total, 512, num_sms() * 16), 256, 0, s>>>(dy, dx, N, D, H, W, sf);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace pcrl
