// GPU-side input staging of the 3-D pre-training path: the intensity / flip / patch-swap /
// normalisation transforms the reference applies per item on the CPU through torchio
// (data.py:73-89, datasets/lunaDataset.py:28-81), as batched kernels over [B][D][H][W] fp32 volumes.
// Random PARAMETERS are drawn by the host (pcrlv2_b200/staging.py) and passed in; the kernels are
// deterministic functions of them, which is what the parity tests pin against oracle/augment_oracle.py.
//   RandomFlip      -> aug_flip_kernel          RandomBlur  -> aug_blur_axis_kernel (scipy gaussian_filter1d
//   RandomNoise + RandomGamma -> aug_noise_gamma_kernel      semantics: truncate 4, reflect, fp64 weights)
//   RandomSwap      -> aug_swap_kernel          ZNormalization -> aug_znorm_kernel
// RandomAffine (SimpleITK resampling in torchio) is not built.  All of these are HBM-bound at 4-12 B/voxel.
#include "common.cuh"

namespace pcrl {

// y[b][d][h][w] = x[b][fd][fh][fw], f* = mirrored index when bit 0 / 1 / 2 of mask[b] is set (axis D / H / W)
__global__ void __launch_bounds__(256)
aug_flip_kernel(const float* __restrict__ x, float* __restrict__ y, const int* __restrict__ mask, int D, int H, int W) {
  const int b = blockIdx.y, vol = D * H * W, mk = mask[b];
  const float* xs = x + (size_t)b * vol;
  float* ys = y + (size_t)b * vol;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < vol; v += gridDim.x * blockDim.x) {
    int w = v % W, h = (v / W) % H, d = v / (W * H);
    if (mk & 1) d = D - 1 - d;
    if (mk & 2) h = H - 1 - h;
    if (mk & 4) w = W - 1 - w;
    ys[v] = xs[((size_t)d * H + h) * W + w];
  }
}

// 1-D Gaussian correlation along one axis with a per-sample sigma: scipy.ndimage.gaussian_filter1d
// (what torchio's RandomBlur calls per axis): radius = int(4*sigma + 0.5), weights exp(-x^2 / (2 sigma^2))
// normalised in fp64, boundary mode 'reflect' (half-sample symmetric), accumulation in fp64, sigma <= 1e-15
// = copy.  axis: 0 = D, 1 = H, 2 = W.
#define AUG_MAX_RADIUS 32
__global__ void __launch_bounds__(256)
aug_blur_axis_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ sigma,
                     int sigma_stride, int axis, int D, int H, int W) {
  __shared__ double wgt[AUG_MAX_RADIUS + 1];
  __shared__ int radius_s;
  const int b = blockIdx.y, vol = D * H * W;
  const double sg = (double)sigma[(size_t)b * sigma_stride + axis];
  if (threadIdx.x == 0) {
    int r = (int)(4.0 * sg + 0.5);
    if (r > AUG_MAX_RADIUS) r = AUG_MAX_RADIUS;
    if (!(sg > 1e-15)) r = -1;            // identity
    radius_s = r;
    if (r >= 0) {
      double sum = 0.0;
      for (int i = 0; i <= r; i++) {
        wgt[i] = exp(-0.5 / (sg * sg) * (double)i * (double)i);
        sum += (i == 0 ? 1.0 : 2.0) * wgt[i];
      }
      for (int i = 0; i <= r; i++) wgt[i] /= sum;
    }
  }
  __syncthreads();
  const int r = radius_s;
  const float* xs = x + (size_t)b * vol;
  float* ys = y + (size_t)b * vol;
  const int n = axis == 0 ? D : (axis == 1 ? H : W);
  const int step = axis == 0 ? H * W : (axis == 1 ? W : 1);
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < vol; v += gridDim.x * blockDim.x) {
    if (r < 0) { ys[v] = xs[v]; continue; }
    const int w = v % W, h = (v / W) % H, d = v / (W * H);
    const int i0 = axis == 0 ? d : (axis == 1 ? h : w);
    const int base = v - i0 * step;
    double acc = wgt[0] * (double)xs[v];
    for (int k = 1; k <= r; k++) {
      int a = i0 - k, c = i0 + k;
      while (a < 0 || a >= n) a = a < 0 ? -a - 1 : 2 * n - 1 - a;      // reflect (d c b a | a b c d | d c b a)
      while (c < 0 || c >= n) c = c < 0 ? -c - 1 : 2 * n - 1 - c;
      acc += wgt[k] * ((double)xs[base + a * step] + (double)xs[base + c * step]);
    }
    ys[v] = (float)acc;
  }
}

// counter-based generator (splitmix64 finaliser on (seed, sample, voxel)) -> standard normal (Box-Muller)
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float normal_from(unsigned long long seed, unsigned long long idx) {
  const unsigned long long r = mix64(seed ^ mix64(idx));
  const float u1 = ((float)(unsigned)(r >> 40) + 0.5f) * (1.f / 16777216.f);     // (0, 1)
  const float u2 = ((float)(unsigned)((r >> 8) & 0xFFFFFF) + 0.5f) * (1.f / 16777216.f);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// RandomNoise then RandomGamma (torchio order in data.py:77-79): t = x + noise_mean + noise_std * n,
// y = sign(t) * |t| ^ exp(log_gamma).  `noise` (nullable) supplies n explicitly (parity tests); otherwise
// n comes from the counter-based generator above (seed, sample, voxel).
__global__ void __launch_bounds__(256)
aug_noise_gamma_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ noise,
                       const float* __restrict__ noise_std, const float* __restrict__ log_gamma,
                       unsigned long long seed, int vol) {
  const int b = blockIdx.y;
  const float sd = noise_std[b], gm = expf(log_gamma[b]);
  const float* xs = x + (size_t)b * vol;
  float* ys = y + (size_t)b * vol;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < vol; v += gridDim.x * blockDim.x) {
    const float n = noise ? noise[(size_t)b * vol + v] : normal_from(seed, (unsigned long long)b * (unsigned)vol + v);
    const float t = xs[v] + sd * n;
    const float a = powf(fabsf(t), gm);
    ys[v] = t < 0.f ? -a : a;
  }
}

// RandomSwap: `iters` sequential swaps of two patches (pd, ph, pw) per sample, in place.  corners [B][iters][6]
// = (d0, h0, w0) of the first and of the second patch.  One block per sample: later swaps see earlier ones.
// torchio extracts BOTH patches, writes the first patch at the second location and then the second patch at
// the first location -- where the two overlap the second write wins; the three phases below keep that order.
#define AUG_SWAP_MAX_PER_THREAD 8
__global__ void __launch_bounds__(256)
aug_swap_kernel(float* __restrict__ x, const int* __restrict__ corners, int iters, int pd, int ph, int pw,
                int D, int H, int W) {
  const int b = blockIdx.x, pvol = pd * ph * pw;
  float* xs = x + (size_t)b * D * H * W;
  for (int it = 0; it < iters; it++) {
    const int* c = corners + ((size_t)b * iters + it) * 6;
    float a[AUG_SWAP_MAX_PER_THREAD], bb[AUG_SWAP_MAX_PER_THREAD];
    int q = 0;
    for (int e = threadIdx.x; e < pvol; e += blockDim.x, q++) {
      const int k = e % pw, j = (e / pw) % ph, i = e / (pw * ph);
      a[q] = xs[((size_t)(c[0] + i) * H + c[1] + j) * W + c[2] + k];
      bb[q] = xs[((size_t)(c[3] + i) * H + c[4] + j) * W + c[5] + k];
    }
    __syncthreads();
    q = 0;
    for (int e = threadIdx.x; e < pvol; e += blockDim.x, q++) {
      const int k = e % pw, j = (e / pw) % ph, i = e / (pw * ph);
      xs[((size_t)(c[3] + i) * H + c[4] + j) * W + c[5] + k] = a[q];
    }
    __syncthreads();
    q = 0;
    for (int e = threadIdx.x; e < pvol; e += blockDim.x, q++) {
      const int k = e % pw, j = (e / pw) % ph, i = e / (pw * ph);
      xs[((size_t)(c[0] + i) * H + c[1] + j) * W + c[2] + k] = bb[q];
    }
    __syncthreads();
  }
}

// ZNormalization: y = (x - mean) / std per volume, std with Bessel's correction (torch.Tensor.std default,
// which torchio uses).  One block per sample, fp64 accumulation.
__global__ void __launch_bounds__(512)
aug_znorm_kernel(const float* __restrict__ x, float* __restrict__ y, int vol) {
  __shared__ double red[2][16];
  __shared__ float mean_s, inv_s;
  const int b = blockIdx.x;
  const float* xs = x + (size_t)b * vol;
  float* ys = y + (size_t)b * vol;
  double s1 = 0.0, s2 = 0.0;
  for (int v = threadIdx.x; v < vol; v += blockDim.x) {
    const double t = (double)xs[v];
    s1 += t;
    s2 += t * t;
  }
  for (int o = 16; o >= 1; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) { a += red[0][i]; q += red[1][i]; }
    const double m = a / vol;
    double var = (q - a * m) / (vol > 1 ? vol - 1 : 1);
    if (var < 0) var = 0;
    mean_s = (float)m;
    inv_s = (float)(1.0 / sqrt(var));
  }
  __syncthreads();
  const float m = mean_s, is = inv_s;
  for (int v = threadIdx.x; v < vol; v += blockDim.x) ys[v] = (xs[v] - m) * is;
}

static inline unsigned vblocks(int vol, int B) {
  long long b = (vol + 255) / 256;
  const long long cap = (num_sms() * 8 + B - 1) / B;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

int aug_flip(const float* x, float* y, const int* mask, int B, int D, int H, int W, cudaStream_t s) {
  aug_flip_kernel<<<dim3(vblocks(D * H * W, B), B), 256, 0, s>>>(x, y, mask, D, H, W);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int aug_blur_axis(const float* x, float* y, const float* sigma, int sigma_stride, int axis, int B, int D, int H,
                  int W, cudaStream_t s) {
  PCRL_REQUIRE(axis >= 0 && axis <= 2, "aug_blur_axis: axis must be 0 (D), 1 (H) or 2 (W)");
  PCRL_REQUIRE(x != y, "aug_blur_axis: not an in-place operation");
  aug_blur_axis_kernel<<<dim3(vblocks(D * H * W, B), B), 256, 0, s>>>(x, y, sigma, sigma_stride, axis, D, H, W);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int aug_noise_gamma(const float* x, float* y, const float* noise, const float* noise_std, const float* log_gamma,
                    unsigned long long seed, int B, int vol, cudaStream_t s) {
  aug_noise_gamma_kernel<<<dim3(vblocks(vol, B), B), 256, 0, s>>>(x, y, noise, noise_std, log_gamma, seed, vol);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int aug_swap(float* x, const int* corners, int iters, int pd, int ph, int pw, int B, int D, int H, int W,
             cudaStream_t s) {
  PCRL_REQUIRE(pd >= 1 && ph >= 1 && pw >= 1 && pd <= D && ph <= H && pw <= W, "aug_swap: patch does not fit the volume");
  PCRL_REQUIRE(pd * ph * pw <= 256 * AUG_SWAP_MAX_PER_THREAD, "aug_swap: patch of %d voxels is too large", pd * ph * pw);
  aug_swap_kernel<<<B, 256, 0, s>>>(x, corners, iters, pd, ph, pw, D, H, W);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int aug_znorm(const float* x, float* y, int B, int vol, cudaStream_t s) {
  aug_znorm_kernel<<<B, 512, 0, s>>>(x, y, vol);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace pcrl

// ------------------------------------------------------------------------------------------------------
// Offline crop generator pieces (luna_preprocess.py:132-137, 213-247): the HU window and the
// 4-deep Python voxel loop that looks, for every voxel, for the first of `len_depth` slices along z whose
// value reaches the lung threshold.
namespace pcrl {

// y = (clip(x, hu_min, hu_max) - hu_min) / (hu_max - hu_min), evaluated in fp64 like numpy does (:133-135)
__global__ void __launch_bounds__(256)
hu_window_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, double hu_min, double hu_max) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = (double)x[i];
    v = v < hu_min ? hu_min : (v > hu_max ? hu_max : v);
    y[i] = (float)(1.0 * (v - hu_min) / (hu_max - hu_min));
  }
}

// crop [X][Y][Z + len_depth - ... >= Z + len_depth - 1] (z fastest, pitch zp): for every (i, j, d < Z):
//   k* = first k < len_depth with crop[i][j][d+k] >= thr;  t = crop[i][j][d+k*], dd = k*   (found)
//   t = 0, dd = len_depth - 1                                                             (not found)
// d_img = 1 - dd / (len_depth - 1);  *sum += sum of d_img (the reference's lung-fraction test, :243-247)
__global__ void __launch_bounds__(256)
depth_scan_kernel(const float* __restrict__ crop, float* __restrict__ t_img, float* __restrict__ d_img,
                  double* __restrict__ sum, int X, int Y, int Z, int zp, int len_depth, float thr) {
  __shared__ double red[8];
  const long long n = (long long)X * Y * Z;
  double s = 0.0;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(v % Z);
    const long long ij = v / Z;
    const float* c = crop + ij * zp + d;
    float t = 0.f;
    int dd = len_depth - 1;
    for (int k = 0; k < len_depth; k++)
      if (c[k] >= thr) { t = c[k]; dd = k; break; }
    const float dv = 1.0f - (float)dd / (float)(len_depth - 1);
    t_img[v] = t;
    d_img[v] = dv;
    s += (double)dv;
  }
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) a += red[i];
    atomicAdd(sum, a);
  }
}

int hu_window(const float* x, float* y, long long n, double hu_min, double hu_max, cudaStream_t s) {
  PCRL_REQUIRE(hu_max > hu_min && n > 0, "hu_window: bad arguments");
  long long b = (n + 255) / 256;
  if (b > num_sms() * 16) b = num_sms() * 16;
  hu_window_kernel<<<(unsigned)b, 256, 0, s>>>(x, y, n, hu_min, hu_max);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int depth_scan(const float* crop, float* t_img, float* d_img, double* sum, int X, int Y, int Z, int zp,
               int len_depth, float thr, cudaStream_t s) {
  PCRL_REQUIRE(len_depth >= 2 && zp >= Z + len_depth - 1, "depth_scan: the crop must carry len_depth - 1 extra slices");
  long long b = ((long long)X * Y * Z + 255) / 256;
  if (b > num_sms() * 16) b = num_sms() * 16;
  depth_scan_kernel<<<(unsigned)b, 256, 0, s>>>(crop, t_img, d_img, sum, X, Y, Z, zp, len_depth, thr);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace pcrl
