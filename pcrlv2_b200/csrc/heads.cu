// Single-output-channel heads of the decoder and the fused SGD step.
//
//   deep-supervision head conv : Conv3d(C -> 1, k=3, p=1)   (models/pcrlv2_model_3d.py:60,71)
//   output transition          : Conv3d(64 -> 1, k=1)        (models/pcrlv2_model_3d.py:78)
//   torch.optim.SGD(momentum, weight_decay)                  (train_3d.py:48-51,151)
//
// The head convolutions are GEMV-shaped (N = 1) and HBM/L2-bound: SIMT kernels, 8 channels
// (16 bytes) per thread, shuffle reduction over the channel groups of a voxel.  Inputs are
// H-padded NDHWC bf16 activations; the 1-channel outputs are plain fp32 [N][D][H][W].
#include "common.cuh"

namespace pcrl {

__device__ __forceinline__ void unpack8h(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8h(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// y1[v] = b3 + sum_tap sum_c a[v+tap][c] * w3[tap][c];   y0[v] = b1 + sum_c a[v][c] * w1[c] (optional)
// w3 is [27][C] fp32 (tap-major), w1 is [C].
__global__ void __launch_bounds__(256)
head_fwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ w3, const float* __restrict__ b3,
                const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ y1,
                float* __restrict__ y0, int N, int D, int H, int W, int C) {
  extern __shared__ float ws[];  // [27][C] (+ [C])
  const int C8 = C >> 3;
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) ws[i] = w3[i];
  if (w1) for (int i = threadIdx.x; i < C; i += blockDim.x) ws[27 * C + i] = w1[i];
  __syncthreads();
  const long long total = (long long)N * D * H * W * C8;
  const int c8 = threadIdx.x % C8;
  // block-uniform loop bound: every lane takes part in the shuffles, tail lanes are clamped
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total;
       base += (long long)gridDim.x * blockDim.x) {
    const long long it = base + threadIdx.x;
    const bool live = it < total;
    const long long v = (live ? it : total - 1) / C8;
    const int wq = (int)(v % W);
    const int h = (int)((v / W) % H);
    const int d = (int)((v / ((long long)W * H)) % D);
    const long long n = v / ((long long)W * H * D);
    float acc = 0.f, acc0 = 0.f;
#pragma unroll
    for (int kz = 0; kz < 3; kz++) {
      const int zz = d + kz - 1;
      if (zz < 0 || zz >= D) continue;
#pragma unroll
      for (int ky = 0; ky < 3; ky++) {
        const int yy = h + ky - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
          const int xx = wq + kx - 1;
          if (xx < 0 || xx >= W) continue;
          float f[8];
          unpack8h(*reinterpret_cast<const uint4*>(
                       a + ((((size_t)n * D + zz) * (H + 1) + yy + 1) * W + xx) * C + c8 * 8), f);
          const float* wt = ws + ((kz * 3 + ky) * 3 + kx) * C + c8 * 8;
#pragma unroll
          for (int q = 0; q < 8; q++) acc = fmaf(f[q], wt[q], acc);
          if (w1 && kz == 1 && ky == 1 && kx == 1) {
            const float* w1s = ws + 27 * C + c8 * 8;
#pragma unroll
            for (int q = 0; q < 8; q++) acc0 = fmaf(f[q], w1s[q], acc0);
          }
        }
      }
    }
    for (int s = C8 >> 1; s >= 1; s >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, s);
      acc0 += __shfl_xor_sync(0xffffffffu, acc0, s);
    }
    if (c8 == 0 && live) {
      y1[v] = acc + b3[0];
      if (w1) y0[v] = acc0 + b1[0];
    }
  }
}

// da[v][c] = sum_tap dy1[v - tap] * w3[tap][c]  (+ dy0[v] * w1[c]);  H-padded bf16 output
__global__ void __launch_bounds__(256)
head_bwd_data_kernel(const float* __restrict__ dy1, const float* __restrict__ w3,
                     const float* __restrict__ dy0, const float* __restrict__ w1,
                     __nv_bfloat16* __restrict__ da, int N, int D, int H, int W, int C) {
  extern __shared__ float ws[];
  const int C8 = C >> 3;
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) ws[i] = w3[i];
  if (w1) for (int i = threadIdx.x; i < C; i += blockDim.x) ws[27 * C + i] = w1[i];
  __syncthreads();
  const long long total = (long long)N * D * (H + 1) * W * C8;
  const int c8 = threadIdx.x % C8;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (long long)gridDim.x * blockDim.x) {
    const long long slot = it / C8;
    const int wq = (int)(slot % W);
    const int hp = (int)((slot / W) % (H + 1));
    const int d = (int)((slot / ((long long)W * (H + 1))) % D);
    const long long n = slot / ((long long)W * (H + 1) * D);
    float out[8];
#pragma unroll
    for (int q = 0; q < 8; q++) out[q] = 0.f;
    if (hp >= 1) {
      const int h = hp - 1;
      const float* g = dy1 + (size_t)n * D * H * W;
#pragma unroll
      for (int kz = 0; kz < 3; kz++) {
        const int zz = d - (kz - 1);
        if (zz < 0 || zz >= D) continue;
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
          const int yy = h - (ky - 1);
          if (yy < 0 || yy >= H) continue;
#pragma unroll
          for (int kx = 0; kx < 3; kx++) {
            const int xx = wq - (kx - 1);
            if (xx < 0 || xx >= W) continue;
            const float gv = __ldg(&g[((size_t)zz * H + yy) * W + xx]);
            const float* wt = ws + ((kz * 3 + ky) * 3 + kx) * C + c8 * 8;
#pragma unroll
            for (int q = 0; q < 8; q++) out[q] = fmaf(gv, wt[q], out[q]);
          }
        }
      }
      if (w1) {
        const float g0 = __ldg(&dy0[(((size_t)n * D + d) * H + h) * W + wq]);
        const float* w1s = ws + 27 * C + c8 * 8;
#pragma unroll
        for (int q = 0; q < 8; q++) out[q] = fmaf(g0, w1s[q], out[q]);
      }
    }
    *reinterpret_cast<uint4*>(da + (size_t)slot * C + c8 * 8) = pack8h(out);
  }
}

// dw3[tap][c] += sum_u a[u][c] * dy1[u - tap];  dw1[c] += sum_u a[u][c] * dy0[u];
// thread = (voxel run, channel group, dz plane of taps): 9 x 8 accumulators.
__global__ void __launch_bounds__(256)
head_bwd_weight_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ dy1,
                       const float* __restrict__ dy0, float* __restrict__ dw3, float* __restrict__ dw1,
                       int N, int D, int H, int W, int C, int vox_per_thread) {
  extern __shared__ float red[];  // [27][C] + [C]
  const int C8 = C >> 3;
  for (int i = threadIdx.x; i < 28 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int c8 = threadIdx.x % C8;
  const int kz = (threadIdx.x / C8) % 3;
  const int sub = threadIdx.x / (3 * C8);           // voxel lane inside the block
  const int lanes = blockDim.x / (3 * C8);
  const long long total = (long long)N * D * H * W;
  float acc[9][8], acc1[8];
#pragma unroll
  for (int t = 0; t < 9; t++)
#pragma unroll
    for (int q = 0; q < 8; q++) acc[t][q] = 0.f;
#pragma unroll
  for (int q = 0; q < 8; q++) acc1[q] = 0.f;
  if (sub < lanes) {
    const long long base = ((long long)blockIdx.x * lanes + sub) * vox_per_thread;
    for (long long u = base; u < base + vox_per_thread && u < total; u++) {
      const int wq = (int)(u % W);
      const int h = (int)((u / W) % H);
      const int d = (int)((u / ((long long)W * H)) % D);
      const long long n = u / ((long long)W * H * D);
      float f[8];
      unpack8h(*reinterpret_cast<const uint4*>(
                   a + ((((size_t)n * D + d) * (H + 1) + h + 1) * W + wq) * C + c8 * 8), f);
      const float* g = dy1 + (size_t)n * D * H * W;
      const int zz = d - (kz - 1);
      if (zz >= 0 && zz < D) {
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
          const int yy = h - (ky - 1);
#pragma unroll
          for (int kx = 0; kx < 3; kx++) {
            const int xx = wq - (kx - 1);
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            const float gv = in ? __ldg(&g[((size_t)zz * H + yy) * W + xx]) : 0.f;
#pragma unroll
            for (int q = 0; q < 8; q++) acc[ky * 3 + kx][q] = fmaf(gv, f[q], acc[ky * 3 + kx][q]);
          }
        }
      }
      if (dy0 && kz == 1) {
        const float g0 = __ldg(&dy0[(((size_t)n * D + d) * H + h) * W + wq]);
#pragma unroll
        for (int q = 0; q < 8; q++) acc1[q] = fmaf(g0, f[q], acc1[q]);
      }
    }
#pragma unroll
    for (int t = 0; t < 9; t++)
#pragma unroll
      for (int q = 0; q < 8; q++) atomicAdd(&red[(kz * 9 + t) * C + c8 * 8 + q], acc[t][q]);
    if (dy0 && kz == 1)
#pragma unroll
      for (int q = 0; q < 8; q++) atomicAdd(&red[27 * C + c8 * 8 + q], acc1[q]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) atomicAdd(&dw3[i], red[i]);
  if (dy0) for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&dw1[i], red[27 * C + i]);
}

// ------------------------------------------------------------------------------ fused SGD
// One launch updates every parameter segment of the flat fp32 buffers:
//   d = g + wd*p;  buf = first ? d : mu*buf + d;  p -= lr*buf
// Segments whose `active` flag is 0 are skipped entirely (a parameter that received no gradient
// this step has grad None in the reference and torch.optim.SGD skips it, SURVEY note N3).
// seg_first[i] != 0 means the momentum buffer of segment i does not exist yet (buf = d).
__global__ void __launch_bounds__(256)
sgd_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                const long long* __restrict__ seg_off, const int* __restrict__ seg_active,
                const int* __restrict__ seg_first, int nseg, float lr, float mu, float wd,
                float grad_scale) {
  for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
    if (!seg_active[s]) continue;
    const long long b = seg_off[s], e = seg_off[s + 1];
    const bool first = seg_first[s] != 0;
    for (long long i = b + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e;
         i += (long long)gridDim.x * blockDim.x) {
      const float pv = p[i];
      const float d = fmaf(wd, pv, g[i] * grad_scale);
      const float m = first ? d : fmaf(mu, buf[i], d);
      buf[i] = m;
      p[i] = fmaf(-lr, m, pv);
    }
  }
}

// ------------------------------------------------------------------------------ wrappers
int head_fwd(const void* a, const float* w3, const float* b3, const float* w1, const float* b1,
             float* y1, float* y0, int N, int D, int H, int W, int C, cudaStream_t s) {
  PCRL_REQUIRE(C % 8 == 0 && C / 8 <= 32 && ((C / 8) & (C / 8 - 1)) == 0,
               "head_fwd: C=%d must be 8 * a power of two <= 256", C);
  const int C8 = C / 8;
  const long long total = (long long)N * D * H * W * C8;
  long long blocks = (total + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  const size_t smem = (size_t)28 * C * 4;
  head_fwd_kernel<<<(unsigned)blocks, 256, smem, s>>>((const __nv_bfloat16*)a, w3, b3, w1, b1, y1, y0, N, D, H, W, C);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int head_bwd_data(const float* dy1, const float* w3, const float* dy0, const float* w1, void* da,
                  int N, int D, int H, int W, int C, cudaStream_t s) {
  PCRL_REQUIRE(C % 8 == 0 && 256 % (C / 8) == 0, "head_bwd_data: unsupported C=%d", C);
  const int C8 = C / 8;
  const long long total = (long long)N * D * (H + 1) * W * C8;
  long long blocks = (total + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  head_bwd_data_kernel<<<(unsigned)blocks, 256, (size_t)28 * C * 4, s>>>(dy1, w3, dy0, w1, (__nv_bfloat16*)da, N, D, H, W, C);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int head_bwd_weight(const void* a, const float* dy1, const float* dy0, float* dw3, float* dw1,
                    int N, int D, int H, int W, int C, cudaStream_t s) {
  PCRL_REQUIRE(C % 8 == 0 && 3 * (C / 8) <= 256, "head_bwd_weight: unsupported C=%d", C);
  const int C8 = C / 8;
  const int lanes = 256 / (3 * C8);
  const long long total = (long long)N * D * H * W;
  const long long threads_target = (long long)num_sms() * 8 * lanes;
  int vpt = (int)((total + threads_target - 1) / threads_target);
  if (vpt < 1) vpt = 1;
  const long long blocks = (total + (long long)lanes * vpt - 1) / ((long long)lanes * vpt);
  static bool configured = false;
  if (!configured) {
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(head_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    configured = true;
  }
  head_bwd_weight_kernel<<<(unsigned)blocks, 256, (size_t)28 * C * 4, s>>>((const __nv_bfloat16*)a, dy1, dy0, dw3, dw1, N, D, H, W, C, vpt);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int sgd_flat(float* p, const float* g, float* buf, const long long* seg_off, const int* seg_active,
             const int* seg_first, int nseg, float lr, float mu, float wd, float grad_scale,
             cudaStream_t s) {
  dim3 grid(64, nseg < 256 ? nseg : 256);
  sgd_flat_kernel<<<grid, 256, 0, s>>>(p, g, buf, seg_off, seg_active, seg_first, nseg, lr, mu, wd, grad_scale);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace pcrl
