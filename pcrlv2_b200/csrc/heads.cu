// Single-output-channel heads of the decoder, the 1-channel mask normalisation, the stem
// weight-gradient staging and the fused SGD step.
//
//   deep-supervision head conv : Conv3d(C -> 1, k=3, p=1)   (models/pcrlv2_model_3d.py:60,71)
//   output transition          : Conv3d(64 -> 1, k=1)        (models/pcrlv2_model_3d.py:78)
//   BatchNorm3d(1) / InstanceNorm3d(1) + Sigmoid of the head (models/pcrlv2_model_3d.py:12,27)
//   torch.optim.SGD(momentum, weight_decay)                  (train_3d.py:48-51,151)
//
// The N=1 convolutions are factored so that the channel contraction runs on the tensor cores:
//   T[u][tap] = sum_c a[u][c] * w[tap][c]        plain GEMM  [rows x C] x [C x 32]   (27 taps, the
//                                                 1x1x1 output conv as column 27, 4 zero columns)
//   y1[v]     = b + sum_tap T[v + tap][tap]      27-point gather on the 32-column fp32 tensor T
// and the same in reverse for the backward pass (scatter dy1 into dT[u][tap] = dy1[u - tap], then
// dA = dT * W and dW = dT^T * A as GEMMs).  Only the gather / scatter are SIMT kernels; they touch
// 128 bytes per voxel instead of 27 * C * 2.
#include "common.cuh"

namespace pcrl {

typedef __nv_bfloat16 bf16_t;
__device__ __forceinline__ void cvt_store1(bf16_t* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void cvt_store1(float* p, float v) { *p = rna_tf32(v); }
struct f32x { float v; };   // fp32 stored exactly (PCRL_DTYPE_F32X), see streaming.cu
__device__ __forceinline__ void cvt_store1(f32x* p, float v) { p->v = v; }
__device__ __forceinline__ void store_row32(f32x* p, const float (&v)[32]) {
  float4* o = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// store 32 consecutive values of one row
__device__ __forceinline__ void store_row32(bf16_t* p, const float (&v)[32]) {
  uint4* o = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 u;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; j++) hh[j] = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
    o[i] = u;
  }
}
__device__ __forceinline__ void store_row32(float* p, const float (&v)[32]) {
  float4* o = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; i++)
    o[i] = make_float4(rna_tf32(v[4 * i]), rna_tf32(v[4 * i + 1]), rna_tf32(v[4 * i + 2]), rna_tf32(v[4 * i + 3]));
}

// w3 (1,C,3,3,3) fp32 [+ w1 (1,C,1,1,1)] -> wext [32][C] bf16 (rows = taps, row 27 = w1) and
// wextT [C][32] bf16
template <typename T>
__global__ void head_pack_kernel(const float* __restrict__ w3, const float* __restrict__ w1,
                                 T* __restrict__ wext, T* __restrict__ wextT, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 32 * C) return;
  const int tap = i / C, c = i % C;
  float v = 0.f;
  if (tap < 27) v = w3[c * 27 + tap];
  else if (tap == 27 && w1) v = w1[c];
  cvt_store1(&wext[tap * C + c], v);
  cvt_store1(&wextT[c * 32 + tap], v);
}

// tT [32][rows] fp32 (rows = N*D*(H+1)*W, H-padded order) -> y1 [N][D][H][W] (+ y0), and the
// per-group sum / sum of squares of y1 for the 1-channel normalisation (stats [G][2] fp64).
__global__ void __launch_bounds__(256)
head_gather_kernel(const float* __restrict__ tT, const float* __restrict__ b3, const float* __restrict__ b1,
                   float* __restrict__ y1, float* __restrict__ y0, double* __restrict__ stats,
                   int stats_per_sample, long long rows, int N, int D, int H, int W) {
  __shared__ float red[2][8];
  const int n = blockIdx.y;
  const int vol = D * H * W;
  const float bias3 = b3[0], bias1 = b1 ? b1[0] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < vol; v += gridDim.x * blockDim.x) {
    const int w = v % W;
    const int h = (v / W) % H;
    const int d = v / (W * H);
    float acc = bias3;
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
          const int xx = w + kx - 1;
          if (xx < 0 || xx >= W) continue;
          const long long r = (((long long)n * D + zz) * (H + 1) + yy + 1) * W + xx;
          acc += __ldg(&tT[(size_t)((kz * 3 + ky) * 3 + kx) * rows + r]);
        }
      }
    }
    y1[(size_t)n * vol + v] = acc;
    if (y0) {
      const long long r = (((long long)n * D + d) * (H + 1) + h + 1) * W + w;
      y0[(size_t)n * vol + v] = __ldg(&tT[(size_t)27 * rows + r]) + bias1;
    }
    s1 += acc;
    s2 += acc * acc;
  }
  if (stats) {
    for (int o = 16; o >= 1; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[threadIdx.x][i];
      atomicAdd(&stats[(stats_per_sample ? (size_t)n * 2 : 0) + threadIdx.x], (double)t);
    }
  }
}

// dT [rows][32] bf16: dT[u][tap] = dy1[u - tap] (0 outside / on pad rows), dT[u][27] = dy0[u]
template <typename T>
__global__ void __launch_bounds__(256)
head_scatter_kernel(const float* __restrict__ dy1, const float* __restrict__ dy0,
                    T* __restrict__ dT, int N, int D, int H, int W) {
  const long long rows = (long long)N * D * (H + 1) * W;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(r % W);
    const int hp = (int)((r / W) % (H + 1));
    const int d = (int)((r / ((long long)W * (H + 1))) % D);
    const long long n = r / ((long long)W * (H + 1) * D);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = 0.f;
    if (hp >= 1) {
      const int h = hp - 1;
      const float* g = dy1 + (size_t)n * D * H * W;
#pragma unroll
      for (int kz = 0; kz < 3; kz++) {
        const int zz = d - (kz - 1);
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
          const int yy = h - (ky - 1);
#pragma unroll
          for (int kx = 0; kx < 3; kx++) {
            const int xx = w - (kx - 1);
            const bool in = zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W;
            v[(kz * 3 + ky) * 3 + kx] = in ? __ldg(&g[((size_t)zz * H + yy) * W + xx]) : 0.f;
          }
        }
      }
      if (dy0) v[27] = __ldg(&dy0[(((size_t)n * D + d) * H + h) * W + w]);
    }
    store_row32(dT + (size_t)r * 32, v);
  }
}

// x [N][D][H][W] fp32 (C = 1) -> X27 [rows][32] bf16 in H-padded row order:
// X27[u][tap] = x[u + tap] (0 outside, pad rows all zero).  Feeds the stem weight gradient GEMM.
template <typename T>
__global__ void __launch_bounds__(256)
im2col27_kernel(const float* __restrict__ x, T* __restrict__ out, int N, int D, int H, int W) {
  const long long rows = (long long)N * D * (H + 1) * W;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(r % W);
    const int hp = (int)((r / W) % (H + 1));
    const int d = (int)((r / ((long long)W * (H + 1))) % D);
    const long long n = r / ((long long)W * (H + 1) * D);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = 0.f;
    if (hp >= 1) {
      const int h = hp - 1;
      const float* g = x + (size_t)n * D * H * W;
#pragma unroll
      for (int kz = 0; kz < 3; kz++) {
        const int zz = d + kz - 1;
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
          const int yy = h + ky - 1;
#pragma unroll
          for (int kx = 0; kx < 3; kx++) {
            const int xx = w + kx - 1;
            const bool in = zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W;
            v[(kz * 3 + ky) * 3 + kx] = in ? __ldg(&g[((size_t)zz * H + yy) * W + xx]) : 0.f;
          }
        }
      }
    }
    store_row32(out + (size_t)r * 32, v);
  }
}

// ------------------------------------------------------------------------------ 1-channel norm + sigmoid
// mask = sigmoid(y*scale + shift) on fp32 [G][vol] (G groups share nothing: scale/shift per group)
__global__ void __launch_bounds__(256)
chan1_sigmoid_fwd_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                         const float* __restrict__ shift, float* __restrict__ mask, int per_sample,
                         long long vol) {
  const int n = blockIdx.y;
  const float sc = scale[per_sample ? n : 0], sh = shift[per_sample ? n : 0];
  const float* yy = y + (size_t)n * vol;
  float* mm = mask + (size_t)n * vol;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < vol;
       i += (long long)gridDim.x * blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(yy + i);
    float4 o;
    o.x = 1.f / (1.f + expf(-fmaf(v.x, sc, sh)));
    o.y = 1.f / (1.f + expf(-fmaf(v.y, sc, sh)));
    o.z = 1.f / (1.f + expf(-fmaf(v.z, sc, sh)));
    o.w = 1.f / (1.f + expf(-fmaf(v.w, sc, sh)));
    *reinterpret_cast<float4*>(mm + i) = o;
  }
}

// backward of mask = sigmoid(norm(y)): pass 0 accumulates sums[g][3] = (sum dz, sum dz*xhat, 0),
// pass 1 writes dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)); dz = dmask*mask*(1-mask)
__global__ void __launch_bounds__(256)
chan1_sigmoid_bwd_kernel(const float* __restrict__ y, const float* __restrict__ mask,
                         const float* __restrict__ dmask, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                         double* __restrict__ sums, float* __restrict__ dy, double count,
                         int per_sample, int pass, long long vol) {
  __shared__ float red[2][8];
  const int n = blockIdx.y;
  const int g = per_sample ? n : 0;
  const float mu = mean[g], is = invstd[g];
  float k1 = 0.f, k2 = 0.f, gs = 0.f;
  if (pass == 1) {
    k1 = (float)(sums[(size_t)g * 3 + 0] / count);
    k2 = (float)(sums[(size_t)g * 3 + 1] / count);
    gs = gamma[0] * is;
  }
  float s0 = 0.f, s1 = 0.f;
  const size_t base = (size_t)n * vol;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < vol;
       i += (long long)gridDim.x * blockDim.x * 4) {
    const float4 yv = *reinterpret_cast<const float4*>(y + base + i);
    const float4 mv = *reinterpret_cast<const float4*>(mask + base + i);
    const float4 gv = *reinterpret_cast<const float4*>(dmask + base + i);
    const float ya[4] = {yv.x, yv.y, yv.z, yv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float out[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float dz = ga[q] * ma[q] * (1.f - ma[q]);
      const float xh = (ya[q] - mu) * is;
      if (pass == 0) { s0 += dz; s1 += dz * xh; }
      else out[q] = gs * (dz - k1 - xh * k2);
    }
    if (pass == 1) *reinterpret_cast<float4*>(dy + base + i) = make_float4(out[0], out[1], out[2], out[3]);
  }
  if (pass == 0) {
    for (int o = 16; o >= 1; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[threadIdx.x][i];
      atomicAdd(&sums[(size_t)g * 3 + threadIdx.x], (double)t);
    }
  }
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

// The same update with every scalar read from DEVICE memory (a captured CUDA graph of the step must
// not bake in the learning rate of one epoch): hyper = [lr, momentum, weight_decay, grad_scale,
// skip_threshold].  guard (nullable): device scalar holding the (rank-summed) loss; the whole update is
// skipped when guard[0] * grad_scale > skip_threshold -- the reference's `loss > 1000 and epoch > 10`
// test (train_3d.py:140), decided identically on every rank.
__global__ void __launch_bounds__(256)
sgd_flat_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                    const long long* __restrict__ seg_off, const int* __restrict__ seg_active,
                    const int* __restrict__ seg_first, int nseg, const float* __restrict__ hyper,
                    const float* __restrict__ guard) {
  const float lr = hyper[0], mu = hyper[1], wd = hyper[2], grad_scale = hyper[3];
  if (guard && guard[0] * grad_scale > hyper[4]) return;
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
static inline unsigned blocks_for(long long items, int per_block, int cap) {
  long long b = (items + per_block - 1) / per_block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int head_pack_weights(const float* w3, const float* w1, void* wext, void* wextT, int C, int dtype, cudaStream_t s) {
  if (dtype == PCRL_DTYPE_F32X) head_pack_kernel<f32x><<<(32 * C + 255) / 256, 256, 0, s>>>(w3, w1, (f32x*)wext, (f32x*)wextT, C);
  else if (dtype == PCRL_DTYPE_F32) head_pack_kernel<float><<<(32 * C + 255) / 256, 256, 0, s>>>(w3, w1, (float*)wext, (float*)wextT, C);
  else head_pack_kernel<bf16_t><<<(32 * C + 255) / 256, 256, 0, s>>>(w3, w1, (bf16_t*)wext, (bf16_t*)wextT, C);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int head_gather(const float* tT, const float* b3, const float* b1, float* y1, float* y0, double* stats,
                int stats_per_sample, int N, int D, int H, int W, cudaStream_t s) {
  const long long rows = (long long)N * D * (H + 1) * W;
  const int vol = D * H * W;
  dim3 grid(blocks_for(vol, 256, (num_sms() * 8 + N - 1) / N), N);
  head_gather_kernel<<<grid, 256, 0, s>>>(tT, b3, b1, y1, y0, stats, stats_per_sample, rows, N, D, H, W);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int head_scatter(const float* dy1, const float* dy0, void* dT, int N, int D, int H, int W, int dtype, cudaStream_t s) {
  const long long rows = (long long)N * D * (H + 1) * W;
  if (dtype == PCRL_DTYPE_F32X) head_scatter_kernel<f32x><<<blocks_for(rows, 256, num_sms() * 16), 256, 0, s>>>(dy1, dy0, (f32x*)dT, N, D, H, W);
  else if (dtype == PCRL_DTYPE_F32) head_scatter_kernel<float><<<blocks_for(rows, 256, num_sms() * 16), 256, 0, s>>>(dy1, dy0, (float*)dT, N, D, H, W);
  else head_scatter_kernel<bf16_t><<<blocks_for(rows, 256, num_sms() * 16), 256, 0, s>>>(dy1, dy0, (bf16_t*)dT, N, D, H, W);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int im2col27(const float* x, void* out, int N, int D, int H, int W, int dtype, cudaStream_t s) {
  const long long rows = (long long)N * D * (H + 1) * W;
  if (dtype == PCRL_DTYPE_F32X) im2col27_kernel<f32x><<<blocks_for(rows, 256, num_sms() * 16), 256, 0, s>>>(x, (f32x*)out, N, D, H, W);
  else if (dtype == PCRL_DTYPE_F32) im2col27_kernel<float><<<blocks_for(rows, 256, num_sms() * 16), 256, 0, s>>>(x, (float*)out, N, D, H, W);
  else im2col27_kernel<bf16_t><<<blocks_for(rows, 256, num_sms() * 16), 256, 0, s>>>(x, (bf16_t*)out, N, D, H, W);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int chan1_sigmoid_fwd(const float* y, const float* scale, const float* shift, float* mask, int per_sample,
                      int G, long long vol, cudaStream_t s) {
  PCRL_REQUIRE(vol % 4 == 0, "chan1_sigmoid_fwd: vol=%lld must be a multiple of 4", vol);
  dim3 grid(blocks_for(vol / 4, 256, (num_sms() * 8 + G - 1) / G), G);
  chan1_sigmoid_fwd_kernel<<<grid, 256, 0, s>>>(y, scale, shift, mask, per_sample, vol);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int chan1_sigmoid_bwd(const float* y, const float* mask, const float* dmask, const float* mean,
                      const float* invstd, const float* gamma, double* sums, float* dy, double count,
                      int per_sample, int pass, int G, long long vol, cudaStream_t s) {
  PCRL_REQUIRE(vol % 4 == 0, "chan1_sigmoid_bwd: vol=%lld must be a multiple of 4", vol);
  dim3 grid(blocks_for(vol / 4, 256, (num_sms() * 8 + G - 1) / G), G);
  chan1_sigmoid_bwd_kernel<<<grid, 256, 0, s>>>(y, mask, dmask, mean, invstd, gamma, sums, dy, count, per_sample, pass, vol);
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

int sgd_flat_dev(float* p, const float* g, float* buf, const long long* seg_off, const int* seg_active,
                 const int* seg_first, int nseg, const float* hyper, const float* guard, cudaStream_t s) {
  dim3 grid(64, nseg < 256 ? nseg : 256);
  sgd_flat_dev_kernel<<<grid, 256, 0, s>>>(p, g, buf, seg_off, seg_active, seg_first, nseg, hyper, guard);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace pcrl
