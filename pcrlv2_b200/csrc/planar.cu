// HBM-bound kernels of the PCRLv2 2-D path (ResNet-18 U-Net, reference models/pcrlv2_model.py and
// torchvision's resnet the smp encoder wraps).  Activations are H-padded NHWC = the layout of
// common.cuh with D = 1: [N][H+1][W][C], row h' = 0 of every image all zero, bf16 or fp32 storage.
// Convolutions that are not 3x3x3 volumes run as  im2col -> tensor-core GEMM (igemm_kmajor.cu /
// igemm_mnmajor.cu plain modes) -> col2im;  everything here moves bytes:
//   im2col2d / col2im2d          any k, stride, padding (7x7/2 stem, 3x3/1, 3x3/2, 1x1/2)
//   maxpool 3x3/2 (+ backward, torch's first-maximum tie rule)
//   add + ReLU (residual join of BasicBlock) and its backward
//   nearest x2 upsampling (+ backward), bilinear upsampling of the 3-channel masks (+ backward)
//   the 3-output-channel convolutions (deep-supervision 1x1, segmentation head 3x3) and their gradients
#include "common.cuh"

namespace pcrl {

typedef __nv_bfloat16 bf16_t;

// ---- 8 consecutive channels <-> registers, templated on the storage type
__device__ __forceinline__ void load8(const bf16_t* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// activations are tensor-core operands of the next GEMM: fp32 storage holds tf32-rounded values
__device__ __forceinline__ void store8(bf16_t* p, const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(rna_tf32(f[0]), rna_tf32(f[1]), rna_tf32(f[2]), rna_tf32(f[3]));
  reinterpret_cast<float4*>(p)[1] = make_float4(rna_tf32(f[4]), rna_tf32(f[5]), rna_tf32(f[6]), rna_tf32(f[7]));
}
__device__ __forceinline__ void zero8(bf16_t* p) { *reinterpret_cast<uint4*>(p) = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ void zero8(float* p) {
  reinterpret_cast<uint4*>(p)[0] = make_uint4(0, 0, 0, 0);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void copy8(const bf16_t* s, bf16_t* d) {
  *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
}
__device__ __forceinline__ void copy8(const float* s, float* d) {
  reinterpret_cast<uint4*>(d)[0] = reinterpret_cast<const uint4*>(s)[0];
  reinterpret_cast<uint4*>(d)[1] = reinterpret_cast<const uint4*>(s)[1];
}
// fp32 stored exactly (PCRL_DTYPE_F32X, precision='fp32x3'): same memory format as float, no tf32 rounding on store
struct f32x { float v; };
__device__ __forceinline__ void load8(const f32x* p, float (&f)[8]) { load8(reinterpret_cast<const float*>(p), f); }
__device__ __forceinline__ void store8(f32x* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void zero8(f32x* p) { zero8(reinterpret_cast<float*>(p)); }
__device__ __forceinline__ void copy8(const f32x* s, f32x* d) {
  copy8(reinterpret_cast<const float*>(s), reinterpret_cast<float*>(d));
}
__device__ __forceinline__ float to_float(f32x v) { return v.v; }
__device__ __forceinline__ void store1(f32x* p, float v) { p->v = v; }
__device__ __forceinline__ float to_float(bf16_t v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ void store1(bf16_t* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void store1(float* p, float v) { *p = rna_tf32(v); }

static inline unsigned grid_for(long long items, int per_block) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ------------------------------------------------------------------------------ im2col / col2im
// col[(n, ho', wo)][(ky*k + kx)*C + c] = x[n][ho*s - p + ky][wo*s - p + kx][c]   (zero outside, zero for ho' = 0
// and for the K-padding columns kk*C .. Kp-1).  C % 8 == 0.
template <typename T>
__global__ void __launch_bounds__(256)
im2col2d_kernel(const T* __restrict__ x, T* __restrict__ col, int N, int H, int W, int C, int k, int s, int p,
                int Ho, int Wo, int Kp) {
  const int kp8 = Kp >> 3, c8n = C >> 3, kk8 = k * k * c8n;
  const long long total = (long long)N * (Ho + 1) * Wo * kp8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % kp8);
    const long long row = i / kp8;
    T* dst = col + row * Kp + (size_t)j * 8;
    const int wo = (int)(row % Wo);
    const int hop = (int)((row / Wo) % (Ho + 1));
    const long long n = row / ((long long)Wo * (Ho + 1));
    bool ok = hop >= 1 && j < kk8;
    int iy = 0, ix = 0, c8 = 0;
    if (ok) {
      const int tap = j / c8n;
      c8 = j - tap * c8n;
      const int ky = tap / k, kx = tap - ky * k;
      iy = (hop - 1) * s - p + ky;
      ix = wo * s - p + kx;
      ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
    }
    if (ok) copy8(x + (((size_t)n * (H + 1) + iy + 1) * W + ix) * C + (size_t)c8 * 8, dst);
    else zero8(dst);
  }
}

// The network input: NCHW fp32 [N][C][H][W], any C (3): same column order, rounded to the operand type.
template <typename T>
__global__ void __launch_bounds__(256)
im2col2d_image_kernel(const float* __restrict__ x, T* __restrict__ col, int N, int H, int W, int C, int k, int s,
                      int p, int Ho, int Wo, int Kp) {
  const int kp8 = Kp >> 3, kk = k * k * C;
  const long long total = (long long)N * (Ho + 1) * Wo * kp8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % kp8);
    const long long row = i / kp8;
    const int wo = (int)(row % Wo);
    const int hop = (int)((row / Wo) % (Ho + 1));
    const long long n = row / ((long long)Wo * (Ho + 1));
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int q = j * 8 + e;
      float t = 0.f;
      if (hop >= 1 && q < kk) {
        const int tap = q / C, c = q - tap * C;
        const int ky = tap / k, kx = tap - ky * k;
        const int iy = (hop - 1) * s - p + ky, ix = wo * s - p + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) t = __ldg(&x[(((size_t)n * C + c) * H + iy) * W + ix]);
      }
      v[e] = t;
    }
    store8(col + row * Kp + (size_t)j * 8, v);
  }
}

// dx[n][iy+1][ix][c] = sum over taps (ky,kx) with (iy + p - ky) % s == 0, (ix + p - kx) % s == 0 of
// dcol[(n, oy+1, ox)][(ky*k+kx)*C + c]  -- the gather form of col2im (no atomics); pad row zero.
template <typename T>
__global__ void __launch_bounds__(256)
col2im2d_kernel(const T* __restrict__ dcol, T* __restrict__ dx, int N, int H, int W, int C, int k, int s, int p,
                int Ho, int Wo, int Kp) {
  const int c8n = C >> 3;
  const long long total = (long long)N * (H + 1) * W * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const long long pix = i / c8n;
    const int ix = (int)(pix % W);
    const int hp = (int)((pix / W) % (H + 1));
    const long long n = pix / ((long long)W * (H + 1));
    T* dst = dx + pix * C + (size_t)c8 * 8;
    if (hp == 0) { zero8(dst); continue; }
    const int iy = hp - 1;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; e++) acc[e] = 0.f;
    for (int ky = 0; ky < k; ky++) {
      const int ty = iy + p - ky;
      if (ty < 0 || ty % s) continue;
      const int oy = ty / s;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < k; kx++) {
        const int tx = ix + p - kx;
        if (tx < 0 || tx % s) continue;
        const int ox = tx / s;
        if (ox >= Wo) continue;
        float v[8];
        load8(dcol + (((size_t)n * (Ho + 1) + oy + 1) * Wo + ox) * Kp + (size_t)(ky * k + kx) * C + (size_t)c8 * 8, v);
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] += v[e];
      }
    }
    store8(dst, acc);
  }
}

// ------------------------------------------------------------------------------ max-pool 3x3 / 2, padding 1
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo) {
  const int c8n = C >> 3;
  const long long total = (long long)N * (Ho + 1) * Wo * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const long long pix = i / c8n;
    const int ox = (int)(pix % Wo);
    const int hp = (int)((pix / Wo) % (Ho + 1));
    const long long n = pix / ((long long)Wo * (Ho + 1));
    T* dst = y + pix * C + (size_t)c8 * 8;
    if (hp == 0) { zero8(dst); continue; }
    const int oy = hp - 1;
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; e++) m[e] = -INFINITY;
    for (int ky = 0; ky < 3; ky++) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 3; kx++) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        load8(x + (((size_t)n * (H + 1) + iy + 1) * W + ix) * C + (size_t)c8 * 8, v);
#pragma unroll
        for (int e = 0; e < 8; e++) m[e] = fmaxf(m[e], v[e]);
      }
    }
    // values are copies of stored inputs: storing them again is exact in both storage types
    store8(dst, m);
  }
}

// dx[pixel] = sum of dy over the (<= 4) windows whose FIRST maximum (scan order ky, kx ascending, strict >:
// torch's rule) is this pixel.  Gather form: every input pixel re-derives the argmax of the windows containing it.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3s2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W,
                      int C, int Ho, int Wo) {
  const int c8n = C >> 3;
  const long long total = (long long)N * (H + 1) * W * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const long long pix = i / c8n;
    const int ix = (int)(pix % W);
    const int hp = (int)((pix / W) % (H + 1));
    const long long n = pix / ((long long)W * (H + 1));
    T* dst = dx + pix * C + (size_t)c8 * 8;
    if (hp == 0) { zero8(dst); continue; }
    const int iy = hp - 1;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; e++) acc[e] = 0.f;
    const T* xn = x + (size_t)n * (H + 1) * W * C + (size_t)c8 * 8;
    // windows (oy, ox) with oy*2-1 <= iy <= oy*2+1
    for (int oy = (iy >> 1); oy <= ((iy + 1) >> 1); oy++) {
      if (oy < 0 || oy >= Ho) continue;
      for (int ox = (ix >> 1); ox <= ((ix + 1) >> 1); ox++) {
        if (ox < 0 || ox >= Wo) continue;
        float m[8];
        int arg[8];
#pragma unroll
        for (int e = 0; e < 8; e++) { m[e] = -INFINITY; arg[e] = -1; }
        for (int ky = 0; ky < 3; ky++) {
          const int yy = oy * 2 - 1 + ky;
          if (yy < 0 || yy >= H) continue;
          for (int kx = 0; kx < 3; kx++) {
            const int xx = ox * 2 - 1 + kx;
            if (xx < 0 || xx >= W) continue;
            float v[8];
            load8(xn + ((size_t)(yy + 1) * W + xx) * C, v);
#pragma unroll
            for (int e = 0; e < 8; e++)
              if (v[e] > m[e] || arg[e] < 0) { m[e] = v[e]; arg[e] = ky * 3 + kx; }
          }
        }
        const int mine = (iy - (oy * 2 - 1)) * 3 + (ix - (ox * 2 - 1));
        float g[8];
        load8(dy + (((size_t)n * (Ho + 1) + oy + 1) * Wo + ox) * C + (size_t)c8 * 8, g);
#pragma unroll
        for (int e = 0; e < 8; e++)
          if (arg[e] == mine) acc[e] += g[e];
      }
    }
    store8(dst, acc);
  }
}

// ------------------------------------------------------------------------------ residual join
template <typename T>
__global__ void __launch_bounds__(256)
add_relu_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float u[8], v[8];
    load8(a + i * 8, u);
    load8(b + i * 8, v);
#pragma unroll
    for (int e = 0; e < 8; e++) u[e] = fmaxf(u[e] + v[e], 0.f);
    store8(out + i * 8, u);
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
add_relu_bwd_kernel(const T* __restrict__ out, const T* __restrict__ g, T* __restrict__ dg, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float o[8], v[8];
    load8(out + i * 8, o);
    load8(g + i * 8, v);
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = o[e] > 0.f ? v[e] : 0.f;
    store8(dg + i * 8, v);
  }
}
// g = a + b (gradient fan-in of a tensor with two consumers)
template <typename T>
__global__ void __launch_bounds__(256)
add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float u[8], v[8];
    load8(a + i * 8, u);
    load8(b + i * 8, v);
#pragma unroll
    for (int e = 0; e < 8; e++) u[e] += v[e];
    store8(out + i * 8, u);
  }
}

// ------------------------------------------------------------------------------ nearest x2
template <typename T>
__global__ void __launch_bounds__(256)
up_nearest2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
  const int c8n = C >> 3, Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)N * (Ho + 1) * Wo * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const long long pix = i / c8n;
    const int ox = (int)(pix % Wo);
    const int hp = (int)((pix / Wo) % (Ho + 1));
    const long long n = pix / ((long long)Wo * (Ho + 1));
    T* dst = y + pix * C + (size_t)c8 * 8;
    if (hp == 0) { zero8(dst); continue; }
    copy8(x + (((size_t)n * (H + 1) + ((hp - 1) >> 1) + 1) * W + (ox >> 1)) * C + (size_t)c8 * 8, dst);
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
up_nearest2_bwd_kernel(const T* __restrict__ g, T* __restrict__ dx, int N, int H, int W, int C) {
  const int c8n = C >> 3, Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)N * (H + 1) * W * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const long long pix = i / c8n;
    const int ix = (int)(pix % W);
    const int hp = (int)((pix / W) % (H + 1));
    const long long n = pix / ((long long)W * (H + 1));
    T* dst = dx + pix * C + (size_t)c8 * 8;
    if (hp == 0) { zero8(dst); continue; }
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; e++) acc[e] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dxx = 0; dxx < 2; dxx++) {
        float v[8];
        load8(g + (((size_t)n * (Ho + 1) + 2 * (hp - 1) + dy + 1) * Wo + 2 * ix + dxx) * C + (size_t)c8 * 8, v);
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] += v[e];
      }
    store8(dst, acc);
  }
}

// ------------------------------------------------------------------------------ bilinear (align_corners=False)
__device__ __forceinline__ void bilin_src(int o, float rscale, int in, int& i0, int& i1, float& l1) {
  float src = ((float)o + 0.5f) * rscale - 0.5f;      // torch: area_pixel_compute_source_index, cubic = false
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - (float)i0;
}
// x fp32 [NC][H][W] -> y fp32 [NC][H*sf][W*sf]
__global__ void __launch_bounds__(256)
bilinear2d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int NC, int H, int W, int sf) {
  const int Ho = H * sf, Wo = W * sf;
  const float rs = 1.f / (float)sf;
  const long long total = (long long)NC * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const long long nc = i / ((long long)Wo * Ho);
    int y0, y1, x0, x1;
    float ly, lx;
    bilin_src(oy, rs, H, y0, y1, ly);
    bilin_src(ox, rs, W, x0, x1, lx);
    const float* p = x + (size_t)nc * H * W;
    const float hy = 1.f - ly, hx = 1.f - lx;
    y[i] = hy * (hx * p[(size_t)y0 * W + x0] + lx * p[(size_t)y0 * W + x1]) +
           ly * (hx * p[(size_t)y1 * W + x0] + lx * p[(size_t)y1 * W + x1]);
  }
}
// dx (zeroed by the caller) += transpose of the above
__global__ void __launch_bounds__(256)
bilinear2d_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, int NC, int H, int W, int sf) {
  const int Ho = H * sf, Wo = W * sf;
  const float rs = 1.f / (float)sf;
  const long long total = (long long)NC * Ho * Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const long long nc = i / ((long long)Wo * Ho);
    int y0, y1, x0, x1;
    float ly, lx;
    bilin_src(oy, rs, H, y0, y1, ly);
    bilin_src(ox, rs, W, x0, x1, lx);
    float* p = dx + (size_t)nc * H * W;
    const float hy = 1.f - ly, hx = 1.f - lx, v = g[i];
    atomicAdd(&p[(size_t)y0 * W + x0], hy * hx * v);
    atomicAdd(&p[(size_t)y0 * W + x1], hy * lx * v);
    atomicAdd(&p[(size_t)y1 * W + x0], ly * hx * v);
    atomicAdd(&p[(size_t)y1 * W + x1], ly * lx * v);
  }
}

// ------------------------------------------------------------------------------ convolutions with 3 output channels
// out[n][j][y][x] = bias[j] + sum_{ky,kx,c} a[n][y+ky-p+1][x+kx-p][c] * w[j][c][ky][kx]     (k = 1 or 3, p = k/2)
// a: [N][H+1][W][Cs] storage type (C <= Cs channels used), w fp32 in the state_dict layout [3][C][k][k],
// out fp32 NCHW.  One thread per pixel; the weights sit in shared memory as [tap][c][3].
template <typename T>
__global__ void __launch_bounds__(128)
conv_c3_fwd_kernel(const T* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out, int N, int H, int W, int C, int Cs, int k) {
  extern __shared__ float ws[];
  const int kk = k * k, p = k >> 1;
  for (int i = threadIdx.x; i < kk * C * 3; i += blockDim.x) {
    const int j = i % 3, c = (i / 3) % C, tap = i / (3 * C);
    ws[i] = w[((size_t)j * C + c) * kk + tap];
  }
  __syncthreads();
  const long long total = (long long)N * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long n = i / ((long long)W * H);
    float acc0 = bias[0], acc1 = bias[1], acc2 = bias[2];
    for (int ky = 0; ky < k; ky++) {
      const int yy = y + ky - p;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < k; kx++) {
        const int xx = x + kx - p;
        if (xx < 0 || xx >= W) continue;
        const T* src = a + (((size_t)n * (H + 1) + yy + 1) * W + xx) * Cs;
        const float* wt = ws + (size_t)(ky * k + kx) * C * 3;
        for (int c = 0; c < C; c += 8) {
          float v[8];
          load8(src + c, v);
#pragma unroll
          for (int e = 0; e < 8; e++) {
            acc0 = fmaf(v[e], wt[(c + e) * 3 + 0], acc0);
            acc1 = fmaf(v[e], wt[(c + e) * 3 + 1], acc1);
            acc2 = fmaf(v[e], wt[(c + e) * 3 + 2], acc2);
          }
        }
      }
    }
    const size_t plane = (size_t)H * W, o = (size_t)n * 3 * plane + (size_t)y * W + x;
    out[o] = acc0;
    out[o + plane] = acc1;
    out[o + 2 * plane] = acc2;
  }
}

// da[n][y+1][x][c] = sum_{j,ky,kx} dout[n][j][y-ky+p][x-kx+p] * w[j][c][ky][kx];  channels C..Cs-1 and the pad row: 0
template <typename T>
__global__ void __launch_bounds__(256)
conv_c3_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ w, T* __restrict__ da, int N,
                        int H, int W, int C, int Cs, int k) {
  extern __shared__ float ws[];           // [tap][j][c]
  const int kk = k * k, p = k >> 1;
  for (int i = threadIdx.x; i < kk * C * 3; i += blockDim.x) {
    const int c = i % C, j = (i / C) % 3, tap = i / (3 * C);
    ws[i] = w[((size_t)j * C + c) * kk + tap];
  }
  __syncthreads();
  const int c8n = Cs >> 3;
  const long long total = (long long)N * (H + 1) * W * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const long long pix = i / c8n;
    const int x = (int)(pix % W);
    const int hp = (int)((pix / W) % (H + 1));
    const long long n = pix / ((long long)W * (H + 1));
    T* dst = da + pix * Cs + (size_t)c8 * 8;
    if (hp == 0 || c8 * 8 >= C) { zero8(dst); continue; }
    const int y = hp - 1;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; e++) acc[e] = 0.f;
    const size_t plane = (size_t)H * W;
    for (int ky = 0; ky < k; ky++) {
      const int yy = y - ky + p;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < k; kx++) {
        const int xx = x - kx + p;
        if (xx < 0 || xx >= W) continue;
        const float* g = dout + (size_t)n * 3 * plane + (size_t)yy * W + xx;
        const float* wt = ws + (size_t)(ky * k + kx) * 3 * C + c8 * 8;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const float gv = __ldg(g + j * plane);
#pragma unroll
          for (int e = 0; e < 8; e++) acc[e] = fmaf(gv, wt[j * C + e], acc[e]);
        }
      }
    }
    store8(dst, acc);
  }
}

// dw[j][c][ky][kx] += sum_{n,y,x} dout[n][j][y][x] * a[n][y+ky-p+1][x+kx-p][c];  db[j] += sum dout[n][j][y][x].
// Thread t of a block owns (tap, c) = (t / C, t % C) for all three j; a block walks a contiguous chunk of pixels
// and adds its partial sums with fp32 atomics (dw, db zeroed by the caller).  blockDim.x = k*k*C (<= 1024).
template <typename T>
__global__ void conv_c3_bwd_weight_kernel(const T* __restrict__ a, const float* __restrict__ dout,
                                          float* __restrict__ dw, float* __restrict__ db, int N, int H, int W,
                                          int C, int Cs, int k, long long chunk) {
  const int kk = k * k, p = k >> 1;
  const int tap = threadIdx.x / C, c = threadIdx.x - tap * C;
  const int ky = tap / k, kx = tap - ky * k;
  const long long total = (long long)N * H * W;
  const long long lo = blockIdx.x * chunk, hi = (lo + chunk < total) ? lo + chunk : total;
  const size_t plane = (size_t)H * W;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
  for (long long i = lo; i < hi; i++) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long n = i / ((long long)W * H);
    const float* g = dout + (size_t)n * 3 * plane + (size_t)y * W + x;
    const float g0 = __ldg(g), g1 = __ldg(g + plane), g2 = __ldg(g + 2 * plane);
    if (threadIdx.x == 0) { b0 += g0; b1 += g1; b2 += g2; }
    const int yy = y + ky - p, xx = x + kx - p;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float v = to_float(a[(((size_t)n * (H + 1) + yy + 1) * W + xx) * Cs + c]);
    s0 = fmaf(g0, v, s0);
    s1 = fmaf(g1, v, s1);
    s2 = fmaf(g2, v, s2);
  }
  atomicAdd(&dw[((size_t)0 * C + c) * kk + tap], s0);
  atomicAdd(&dw[((size_t)1 * C + c) * kk + tap], s1);
  atomicAdd(&dw[((size_t)2 * C + c) * kk + tap], s2);
  if (threadIdx.x == 0) {
    atomicAdd(&db[0], b0);
    atomicAdd(&db[1], b1);
    atomicAdd(&db[2], b2);
  }
}


// ------------------------------------------------------------------------------ weight layout converters
// nn.Conv2d.weight (Cout, Cin, k, k) fp32 -> wmat [CoutP][Kp] and wt [Kp][CoutP] in the operand type;
// K index = (ky*k + kx)*cs + c (cs = channel stride of the input activation), zero outside.
template <typename T>
__global__ void __launch_bounds__(256)
pack_conv2d_kernel(const float* __restrict__ w, T* __restrict__ wmat, T* __restrict__ wt, int Cout, int Cin, int k,
                   int cs, int CoutP, int Kp) {
  const long long total = (long long)CoutP * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % Kp), co = (int)(i / Kp);
    const int tap = q / cs, c = q - tap * cs;
    float v = 0.f;
    if (co < Cout && tap < k * k && c < Cin) v = w[((size_t)co * Cin + c) * k * k + tap];
    store1(wmat + i, v);
    store1(wt + (size_t)q * CoutP + co, v);
  }
}
// dW as the GEMM left it ([CoutP][Kp], or transposed [Kp][CoutP]) -> (Cout, Cin, k, k) fp32
__global__ void __launch_bounds__(256)
unpack_conv2d_wgrad_kernel(const float* __restrict__ dwm, float* __restrict__ g, int Cout, int Cin, int k, int cs,
                           int CoutP, int Kp, int transposed) {
  const int kk = k * k;
  const long long total = (long long)Cout * Cin * kk;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % kk), c = (int)((i / kk) % Cin), co = (int)(i / ((long long)kk * Cin));
    const int q = tap * cs + c;
    g[i] = transposed ? dwm[(size_t)q * CoutP + co] : dwm[(size_t)co * Kp + q];
  }
}

// ------------------------------------------------------------------------------ launchers
// runs STMT with `T` = the storage type selected by `dtype`
#define PCRL_DISPATCH3(dtype, STMT)                                   \
  do {                                                                \
    if ((dtype) == PCRL_DTYPE_BF16) { typedef bf16_t T; STMT; }       \
    else if ((dtype) == PCRL_DTYPE_F32X) { typedef f32x T; STMT; }    \
    else { typedef float T; STMT; }                                   \
  } while (0)

int im2col2d(const void* x, void* col, int N, int H, int W, int C, int k, int s, int p, int Ho, int Wo, int Kp,
             int image_nchw, int dtype, cudaStream_t st) {
  PCRL_REQUIRE(Kp % 8 == 0 && Kp >= k * k * C, "im2col2d: Kp=%d must be a multiple of 8 and >= k*k*C", Kp);
  PCRL_REQUIRE(Ho == (H + 2 * p - k) / s + 1 && Wo == (W + 2 * p - k) / s + 1, "im2col2d: output size mismatch");
  const long long total = (long long)N * (Ho + 1) * Wo * (Kp / 8);
  const unsigned g = grid_for(total, 256);
  if (image_nchw) {
    PCRL_DISPATCH3(dtype, (im2col2d_image_kernel<T><<<g, 256, 0, st>>>((const float*)x, (T*)col, N, H, W, C, k, s, p, Ho, Wo, Kp)));
  } else {
    PCRL_REQUIRE(C % 8 == 0, "im2col2d: C=%d must be a multiple of 8", C);
    PCRL_DISPATCH3(dtype, (im2col2d_kernel<T><<<g, 256, 0, st>>>((const T*)x, (T*)col, N, H, W, C, k, s, p, Ho, Wo, Kp)));
  }
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int col2im2d(const void* dcol, void* dx, int N, int H, int W, int C, int k, int s, int p, int Ho, int Wo, int Kp,
             int dtype, cudaStream_t st) {
  PCRL_REQUIRE(C % 8 == 0 && Kp >= k * k * C, "col2im2d: C=%d must be a multiple of 8, Kp=%d >= k*k*C", C, Kp);
  const long long total = (long long)N * (H + 1) * W * (C / 8);
  PCRL_DISPATCH3(dtype, (col2im2d_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)dcol, (T*)dx, N, H, W, C, k, s, p, Ho, Wo, Kp)));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int maxpool2d_3x3s2(const void* x, const void* dy, void* out, int N, int H, int W, int C, int backward, int dtype,
                    cudaStream_t st) {
  PCRL_REQUIRE(C % 8 == 0, "maxpool2d: C=%d must be a multiple of 8", C);
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  if (!backward) {
    const long long total = (long long)N * (Ho + 1) * Wo * (C / 8);
    PCRL_DISPATCH3(dtype, (maxpool3s2_fwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)x, (T*)out, N, H, W, C, Ho, Wo)));
  } else {
    const long long total = (long long)N * (H + 1) * W * (C / 8);
    PCRL_DISPATCH3(dtype, (maxpool3s2_bwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)x, (const T*)dy, (T*)out, N, H, W, C, Ho, Wo)));
  }
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// op: 0 = relu(a + b), 1 = backward of it (a = out, b = g), 2 = a + b
int add_relu(const void* a, const void* b, void* out, long long n, int op, int dtype, cudaStream_t st) {
  PCRL_REQUIRE(n % 8 == 0 && op >= 0 && op <= 2, "add_relu: n=%lld must be a multiple of 8, op in 0..2", n);
  const long long n8 = n / 8;
  const unsigned g = grid_for(n8, 256);
  if (op == 0) PCRL_DISPATCH3(dtype, (add_relu_fwd_kernel<T><<<g, 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, n8)));
  else if (op == 1) PCRL_DISPATCH3(dtype, (add_relu_bwd_kernel<T><<<g, 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, n8)));
  else PCRL_DISPATCH3(dtype, (add_kernel<T><<<g, 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, n8)));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int upsample_nearest2x(const void* x, void* y, int N, int H, int W, int C, int backward, int dtype, cudaStream_t st) {
  PCRL_REQUIRE(C % 8 == 0, "upsample_nearest2x: C=%d must be a multiple of 8", C);
  // forward: x coarse [N][H+1][W][C] -> y fine [N][2H+1][2W][C]; backward: x = fine gradient, y = coarse gradient
  if (!backward) {
    const long long total = (long long)N * (2 * H + 1) * 2 * W * (C / 8);
    PCRL_DISPATCH3(dtype, (up_nearest2_fwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C)));
  } else {
    const long long total = (long long)N * (H + 1) * W * (C / 8);
    PCRL_DISPATCH3(dtype, (up_nearest2_bwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C)));
  }
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int bilinear2d(const float* x, float* y, int NC, int H, int W, int sf, int backward, cudaStream_t st) {
  PCRL_REQUIRE(sf >= 1, "bilinear2d: scale factor %d", sf);
  const long long total = (long long)NC * H * sf * W * sf;
  if (!backward) bilinear2d_fwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, y, NC, H, W, sf);
  else bilinear2d_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, y, NC, H, W, sf);   // x = g (fine), y = dx (coarse, zeroed)
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int conv2d_c3_fwd(const void* a, const float* w, const float* bias, float* out, int N, int H, int W, int C, int Cs,
                  int k, int dtype, cudaStream_t st) {
  PCRL_REQUIRE((k == 1 || k == 3) && C % 8 == 0 && Cs % 8 == 0 && C <= Cs, "conv2d_c3: k=%d C=%d Cs=%d", k, C, Cs);
  const size_t smem = (size_t)k * k * C * 3 * sizeof(float);
  PCRL_REQUIRE(smem <= 48 * 1024, "conv2d_c3: k*k*C=%d too large", k * k * C);
  const long long total = (long long)N * H * W;
  PCRL_DISPATCH3(dtype, (conv_c3_fwd_kernel<T><<<grid_for(total, 128), 128, smem, st>>>((const T*)a, w, bias, out, N, H, W, C, Cs, k)));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int conv2d_c3_bwd(const void* a, const float* w, const float* dout, void* da, float* dw, float* db, int N, int H,
                  int W, int C, int Cs, int k, int dtype, cudaStream_t st) {
  PCRL_REQUIRE((k == 1 || k == 3) && C % 8 == 0 && Cs % 8 == 0 && C <= Cs, "conv2d_c3: k=%d C=%d Cs=%d", k, C, Cs);
  const size_t smem = (size_t)k * k * C * 3 * sizeof(float);
  PCRL_REQUIRE(smem <= 48 * 1024 && k * k * C <= 1024, "conv2d_c3: k*k*C=%d too large", k * k * C);
  if (da) {
    const long long total = (long long)N * (H + 1) * W * (Cs / 8);
    PCRL_DISPATCH3(dtype, (conv_c3_bwd_data_kernel<T><<<grid_for(total, 256), 256, smem, st>>>(dout, w, (T*)da, N, H, W, C, Cs, k)));
    PCRL_CHECK_LAUNCH();
  }
  if (dw) {
    PCRL_REQUIRE(db != nullptr, "conv2d_c3_bwd: db is NULL");
    const long long total = (long long)N * H * W;
    long long blocks = (long long)num_sms() * 4;
    if (blocks > total) blocks = total;
    const long long chunk = (total + blocks - 1) / blocks;
    blocks = (total + chunk - 1) / chunk;
    PCRL_DISPATCH3(dtype, (conv_c3_bwd_weight_kernel<T><<<(unsigned)blocks, k * k * C, 0, st>>>((const T*)a, dout, dw, db, N, H, W, C, Cs, k, chunk)));
    PCRL_CHECK_LAUNCH();
  }
  return PCRL_OK;
}

int pack_conv2d_weights(const float* w, void* wmat, void* wt, int Cout, int Cin, int k, int cs, int CoutP, int Kp,
                        int dtype, cudaStream_t st) {
  PCRL_REQUIRE(CoutP >= Cout && cs >= Cin && Kp >= k * k * cs, "pack_conv2d_weights: bad padded dims");
  const long long total = (long long)CoutP * Kp;
  PCRL_DISPATCH3(dtype, (pack_conv2d_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(w, (T*)wmat, (T*)wt, Cout, Cin, k, cs, CoutP, Kp)));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int unpack_conv2d_wgrad(const float* dwm, float* g, int Cout, int Cin, int k, int cs, int CoutP, int Kp, int transposed,
                        cudaStream_t st) {
  const long long total = (long long)Cout * Cin * k * k;
  unpack_conv2d_wgrad_kernel<<<grid_for(total, 256), 256, 0, st>>>(dwm, g, Cout, Cin, k, cs, CoutP, Kp, transposed);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace pcrl
