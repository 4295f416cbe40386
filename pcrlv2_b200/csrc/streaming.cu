// HBM-bound kernels of the PCRLv2 3-D path: weight packers, the Cin=1 stem convolution, the
// norm(+activation)(+2x2x2 max-pool)(+global-average-pool sums) forward pass and its two-pass
// backward, ConvTranspose gradient un-shuffle.  All activations are H-padded NDHWC bf16
// (common.cuh); every producer writes the zero pad row of each plane itself.
//
// Reference ops replaced (models/pcrlv2_model_3d.py): BatchNorm3d / InstanceNorm3d :11-16,
// ReLU / PReLU / ELU / Sigmoid :20-27, MaxPool3d(2) :100,115-117, adaptive_avg_pool3d :67,
// Conv3d(1->32) :114 (down_tr64.ops.0), ConvTranspose3d :52 (backward re-layout).
#include "common.cuh"

namespace pcrl {

enum { ACT_RELU = 0, ACT_PRELU = 1, ACT_ELU = 2, ACT_SIGMOID = 3, ACT_NONE = 4, ACT_LEAKY = 5 };

// ---- 8-channel vector access, templated on the storage type (bf16: 16 bytes, fp32: 32 bytes)
typedef __nv_bfloat16 bf16_t;
template <typename T> struct V8;
template <> struct V8<bf16_t> { uint4 a; };
template <> struct V8<float> { uint4 a, b; };

__device__ __forceinline__ V8<bf16_t> ld8(const bf16_t* p) {
  V8<bf16_t> v;
  v.a = *reinterpret_cast<const uint4*>(p);
  return v;
}
__device__ __forceinline__ V8<float> ld8(const float* p) {
  V8<float> v;
  v.a = reinterpret_cast<const uint4*>(p)[0];
  v.b = reinterpret_cast<const uint4*>(p)[1];
  return v;
}
__device__ __forceinline__ void up8(const V8<bf16_t>& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v.a);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void up8(const V8<float>& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.a.x); f[1] = __uint_as_float(v.a.y); f[2] = __uint_as_float(v.a.z); f[3] = __uint_as_float(v.a.w);
  f[4] = __uint_as_float(v.b.x); f[5] = __uint_as_float(v.b.y); f[6] = __uint_as_float(v.b.z); f[7] = __uint_as_float(v.b.w);
}
__device__ __forceinline__ void st8(bf16_t* p, const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void st8(float* p, const float (&f)[8]) {
  // fp32 activations are tensor-core (tf32) operands of the next kernel: store them tf32-rounded
  reinterpret_cast<float4*>(p)[0] = make_float4(rna_tf32(f[0]), rna_tf32(f[1]), rna_tf32(f[2]), rna_tf32(f[3]));
  reinterpret_cast<float4*>(p)[1] = make_float4(rna_tf32(f[4]), rna_tf32(f[5]), rna_tf32(f[6]), rna_tf32(f[7]));
}
// fp32 stored exactly (PCRL_DTYPE_F32X): same memory format as float, no tf32 rounding on store
struct f32x { float v; };
template <> struct V8<f32x> { uint4 a, b; };
__device__ __forceinline__ V8<f32x> ld8(const f32x* p) {
  V8<f32x> v;
  v.a = reinterpret_cast<const uint4*>(p)[0];
  v.b = reinterpret_cast<const uint4*>(p)[1];
  return v;
}
__device__ __forceinline__ void up8(const V8<f32x>& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.a.x); f[1] = __uint_as_float(v.a.y); f[2] = __uint_as_float(v.a.z); f[3] = __uint_as_float(v.a.w);
  f[4] = __uint_as_float(v.b.x); f[5] = __uint_as_float(v.b.y); f[6] = __uint_as_float(v.b.z); f[7] = __uint_as_float(v.b.w);
}
__device__ __forceinline__ void st8(f32x* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void z8(f32x* p) {
  reinterpret_cast<uint4*>(p)[0] = make_uint4(0, 0, 0, 0);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void rnd8(f32x, float (&)[8]) {}
__device__ __forceinline__ void cvt_store(f32x* p, float v) { p->v = v; }
__device__ __forceinline__ float to_f(f32x v) { return v.v; }
__device__ __forceinline__ void z8(bf16_t* p) { *reinterpret_cast<uint4*>(p) = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ void z8(float* p) {
  reinterpret_cast<uint4*>(p)[0] = make_uint4(0, 0, 0, 0);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(0, 0, 0, 0);
}
// round through the storage type: statistics / sums are taken over the values that are stored
__device__ __forceinline__ void rnd8(bf16_t, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) f[i] = __bfloat162float(__float2bfloat16(f[i]));
}
__device__ __forceinline__ void rnd8(float, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) f[i] = rna_tf32(f[i]);
}
__device__ __forceinline__ void cvt_store(bf16_t* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void cvt_store(float* p, float v) { *p = rna_tf32(v); }
__device__ __forceinline__ float to_f(bf16_t v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(float v) { return v; }

__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
  switch (act) {
    case ACT_RELU: return fmaxf(z, 0.f);
    case ACT_PRELU: return z > 0.f ? z : slope * z;
    case ACT_ELU: return z > 0.f ? z : expm1f(z);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    case ACT_LEAKY: return z > 0.f ? z : 0.01f * z;
    default: return z;
  }
}
// derivative of the activation wrt its input z
__device__ __forceinline__ float act_bwd(float z, int act, float slope) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case ACT_PRELU: return z > 0.f ? 1.f : slope;
    case ACT_ELU: return z > 0.f ? 1.f : expf(z);
    case ACT_SIGMOID: { float s = 1.f / (1.f + expf(-z)); return s * (1.f - s); }
    case ACT_LEAKY: return z > 0.f ? 1.f : 0.01f;
    default: return 1.f;
  }
}

// ------------------------------------------------------------------------------ weight packers
// w [Cout][Cin][27] fp32 (tap = (kz*3+ky)*3+kx) -> tensor-core operand layouts of igemm_kmajor:
//   wf [9 (ky,kx)][3 (kz = 2,1,0)][Cout][Cin] bf16   forward
//   wd [9][3][Cin][Cout] bf16, taps mirrored          data gradient
// (the kz-innermost order lets one TMA box fetch the filters of up to three stacked output planes)
template <typename T>
__global__ void pack_conv3_kernel(const float* __restrict__ w, T* __restrict__ wf,
                                  T* __restrict__ wd, int Cout, int Cin) {
  const long long total = (long long)Cout * Cin * 27;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 27);
    const long long r = i / 27;
    const int ci = (int)(r % Cin), co = (int)(r / Cin);
    const int kz = tap / 9, ky = (tap / 3) % 3, kx = tap % 3;
    const float v = w[i];
    const int tf = (ky * 3 + kx) * 3 + (2 - kz);
    cvt_store(&wf[((size_t)tf * Cout + co) * Cin + ci], v);
    if (wd) {
      const int td = ((2 - ky) * 3 + (2 - kx)) * 3 + kz;
      cvt_store(&wd[((size_t)td * Cin + ci) * Cout + co], v);
    }
  }
}
// gpk [27][Cout][Cin] fp32 -> g [Cout][Cin][27] fp32
__global__ void unpack_conv3_wgrad_kernel(const float* __restrict__ gpk, float* __restrict__ g,
                                          int Cout, int Cin) {
  const long long total = (long long)Cout * Cin * 27;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 27);
    const long long r = i / 27;
    const int ci = (int)(r % Cin), co = (int)(r / Cin);
    g[i] = gpk[((size_t)tap * Cout + co) * Cin + ci];
  }
}
// w [Cin][Cout][8] fp32 -> wf [(t,co)][Cin] bf16 (forward B operand) and wd [Cin][(t,co)] bf16
template <typename T>
__global__ void pack_convT_kernel(const float* __restrict__ w, T* __restrict__ wf,
                                  T* __restrict__ wd, int Cin, int Cout) {
  const long long total = (long long)Cin * Cout * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % 8);
    const long long r = i / 8;
    const int co = (int)(r % Cout), ci = (int)(r / Cout);
    const float v = w[i];
    cvt_store(&wf[((size_t)t * Cout + co) * Cin + ci], v);
    cvt_store(&wd[(size_t)ci * (8 * Cout) + (size_t)t * Cout + co], v);
  }
}
// gpk [(t,co)][Cin] fp32 -> g [Cin][Cout][8]
__global__ void unpack_convT_wgrad_kernel(const float* __restrict__ gpk, float* __restrict__ g,
                                          int Cin, int Cout) {
  const long long total = (long long)Cin * Cout * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % 8);
    const long long r = i / 8;
    const int co = (int)(r % Cout), ci = (int)(r / Cout);
    g[i] = gpk[((size_t)t * Cout + co) * Cin + ci];
  }
}

// ------------------------------------------------------------------------------ stem conv, Cin = 1
// x [N][D][H][W] fp32 (C = 1, so NCDHW == NDHWC), w [32][27] fp32, y H-padded bf16 [N][D][H+1][W][32].
// One thread per output voxel, 32 output channels in registers; statistics as in the igemm epilogue.
template <typename T>
__global__ void __launch_bounds__(128)
stem_conv_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w,
                       T* __restrict__ y, double* __restrict__ stats,
                       int stats_per_sample, int N, int D, int H, int W) {
  __shared__ float ws[27][32];
  __shared__ float st[2][32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) ws[i % 27][i / 27] = w[i];
  if (threadIdx.x < 64) st[threadIdx.x / 32][threadIdx.x % 32] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int H1 = H + 1;
  const int slots = D * H1 * W;  // voxel slots of one sample, pad rows included
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; i++) acc[i] = 0.f;
  bool real = false;
  if (slot < slots) {
    const int wq = slot % W;
    const int hp = (slot / W) % H1;
    const int d = slot / (W * H1);
    real = hp >= 1;
    if (real) {
      const int h = hp - 1;
      const float* xs = x + (size_t)n * D * H * W;
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
            const float v = __ldg(&xs[((size_t)zz * H + yy) * W + xx]);
            const float* wt = ws[(kz * 3 + ky) * 3 + kx];
#pragma unroll
            for (int c = 0; c < 32; c++) acc[c] = fmaf(v, wt[c], acc[c]);
          }
        }
      }
    }
    // store (pad rows get zeros)
    T* o = y + ((size_t)n * slots + slot) * 32;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; j++) f[j] = acc[8 * i + j];
      rnd8(T(), f);
      st8(o + 8 * i, f);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[8 * i + j] = f[j];  // statistics over the stored values
    }
  }
  if (stats) {
    float s1[32], s2[32];
#pragma unroll
    for (int i = 0; i < 32; i++) { s1[i] = real ? acc[i] : 0.f; s2[i] = s1[i] * s1[i]; }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; i++) {
        const float a1 = up ? s1[i] : s1[i + s], k1 = up ? s1[i + s] : s1[i];
        s1[i] = k1 + __shfl_xor_sync(0xffffffffu, a1, s);
        const float a2 = up ? s2[i] : s2[i + s], k2 = up ? s2[i + s] : s2[i];
        s2[i] = k2 + __shfl_xor_sync(0xffffffffu, a2, s);
      }
    }
    atomicAdd(&st[0][lane], s1[0]);
    atomicAdd(&st[1][lane], s2[0]);
    __syncthreads();
    if (threadIdx.x < 32) {
      double* g = stats + (stats_per_sample ? (size_t)n * 64 : 0);
      atomicAdd(&g[threadIdx.x * 2 + 0], (double)st[0][threadIdx.x]);
      atomicAdd(&g[threadIdx.x * 2 + 1], (double)st[1][threadIdx.x]);
    }
  }
}

// dw[32][27] += sum_v dy[v][co] * x[v + tap]; lane = output channel, one warp per run of voxels.
template <typename T>
__global__ void __launch_bounds__(256)
stem_conv_wgrad_kernel(const T* __restrict__ dy, const float* __restrict__ x,
                       float* __restrict__ dw, int N, int D, int H, int W, int vox_per_warp) {
  __shared__ float red[27][32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) red[i / 32][i % 32] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long total = (long long)N * D * H * W;
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; i++) acc[i] = 0.f;
  const long long v0 = gw * vox_per_warp;
  for (long long v = v0; v < v0 + vox_per_warp && v < total; v++) {
    const int wq = (int)(v % W);
    const int h = (int)((v / W) % H);
    const int d = (int)((v / ((long long)W * H)) % D);
    const int n = (int)(v / ((long long)W * H * D));
    const float g = to_f(dy[((((size_t)n * D + d) * (H + 1) + h + 1) * W + wq) * 32 + lane]);
    const float* xs = x + (size_t)n * D * H * W;
#pragma unroll
    for (int kz = 0; kz < 3; kz++) {
      const int zz = d + kz - 1;
#pragma unroll
      for (int ky = 0; ky < 3; ky++) {
        const int yy = h + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
          const int xx = wq + kx - 1;
          const bool in = zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W;
          const float xv = in ? __ldg(&xs[((size_t)zz * H + yy) * W + xx]) : 0.f;
          acc[(kz * 3 + ky) * 3 + kx] = fmaf(g, xv, acc[(kz * 3 + ky) * 3 + kx]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 27; t++) atomicAdd(&red[t][lane], acc[t]);
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) {
    const int t = i / 32, co = i % 32;
    atomicAdd(&dw[co * 27 + t], red[t][co]);
  }
}

// ------------------------------------------------------------------------------ norm finalize
// stats [G][C][2] fp64 (sum, sum of squares over `count` elements per (g, c)); G = 1 for
// BatchNorm, N for InstanceNorm.  Produces scale/shift (y*scale + shift = gamma*xhat + beta) and
// saves mean / invstd for the backward pass.  BatchNorm running statistics follow
// torch.nn.functional.batch_norm(training=True, momentum): biased variance to normalise, unbiased
// into running_var; `conv_bias` is the bias the convolution did NOT add (it cancels under the
// normalisation, SURVEY note N1) and is folded into running_mean.
__global__ void norm_finalize_kernel(const double* __restrict__ stats, double count,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ conv_bias,
                                     float* __restrict__ running_mean, float* __restrict__ running_var,
                                     long long* __restrict__ num_batches_tracked, float momentum,
                                     float eps, float* __restrict__ scale, float* __restrict__ shift,
                                     float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                     int G, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && num_batches_tracked) *num_batches_tracked += 1;
  if (i >= G * C) return;
  const int c = i % C;
  const double m = stats[2 * i] / count;
  double var = stats[2 * i + 1] / count - m * m;
  if (var < 0) var = 0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  scale[i] = sc;
  shift[i] = beta[c] - (float)m * sc;
  mean_out[i] = (float)m;
  invstd_out[i] = invstd;
  if (running_mean) {
    const float mb = (float)m + (conv_bias ? conv_bias[c] : 0.f);
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mb;
    const double unb = count > 1 ? var * count / (count - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

// ------------------------------------------------------------------------------ norm + act forward
// y [N][D][H+1][W][C] bf16 -> a = act(y*scale + shift), written as
//   a_out    full resolution (nullable),
//   pool_out 2x2x2 max-pool, H-padded [N][D/2][H/2+1][W/2][C] (nullable),
//   avg_sum  [N][C] fp32 += sum over voxels of a (nullable; caller zeroes; divide by D*H*W later).
// Work is mapped by merged row: a block owns whole rows of the (pooled) output, a thread owns 8
// channels (16 bytes) of one column of the row and keeps that channel group for the whole kernel,
// so the only integer division in the loop is the pad-row test.  RELU is specialised at compile
// time (ACT = PCRL_ACT_RELU), every other activation takes the generic path (ACT = -1).
template <typename T>
struct NormActFwdParams {
  const T* y;
  const float* scale; const float* shift; const float* prelu;
  T* a_out; T* pool_out; float* avg_sum;
  int per_sample, act, N, D, H, W, C;
  int wseg;          // a row of W voxels is processed as wseg segments (rows wider than one block)
};

struct RowMap {
  int rows_per_iter, t_row, w, c8, items_per_row, chunks;
};
__device__ __forceinline__ RowMap make_row_map(int cw, int C8) {
  RowMap m;
  m.items_per_row = cw * C8;
  m.rows_per_iter = (int)blockDim.x / m.items_per_row;
  if (m.rows_per_iter < 1) m.rows_per_iter = 1;
  m.chunks = (m.items_per_row + (int)blockDim.x - 1) / (int)blockDim.x;  // > 1 only for very wide rows
  const int t_item = (int)threadIdx.x % m.items_per_row;
  m.t_row = (int)threadIdx.x / m.items_per_row;
  m.w = t_item / C8;
  m.c8 = t_item % C8;
  return m;
}

template <int ACT>
__device__ __forceinline__ float act_f(float z, int act, float slope) {
  if (ACT == ACT_RELU) return fmaxf(z, 0.f);
  return act_fwd(z, act, slope);
}
template <int ACT>
__device__ __forceinline__ float act_d(float z, int act, float slope) {
  if (ACT == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return act_bwd(z, act, slope);
}

// Per-channel constants of the norm/act kernels live in SHARED memory, not in registers: a thread owns
// 8 channels, and 5-6 constant arrays of 8 floats each pushed the two-pass backward kernel to 154-255
// registers = ONE resident block per SM (ncu, profiles/r02g_ncu_norm_act.md: 12 % warps active, 2.3 TB/s
// in bf16).  They are fetched per use with volatile shared loads (the compiler would otherwise hoist them
// back into registers); with two rows in flight per thread the kernels fit 3 blocks per SM.
__device__ __forceinline__ void ldc8(const float* p, float (&f)[8]) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(a + 16));
}
// slots 3 / 4 hold (invstd, mean*invstd) in the statistics pass and (gs, c1) in the apply pass: 7 slots keep
// the largest layer (C = 512) under the 48 KB default dynamic shared memory limit
enum { K_SC = 0, K_SH = 1, K_SL = 2, K_IS = 3, K_MIS = 4, K_GS = 3, K_C1 = 4, K_GA = 5, K_C2 = 6, K_COUNT = 7 };
// Layout of one constant array in shared memory: [C/8][12] floats -- the 8 values of a channel group plus 4
// floats of padding.  With a plain [C] array the 32-byte reads of the 8 channel groups of a quarter-warp
// sit 32 bytes apart and collide pairwise in the banks (ncu: 4e7 bank conflicts per launch, short-scoreboard
// the top stall); a 48-byte pitch makes both 16-byte halves conflict-free.
#define CPITCH 12
__device__ __forceinline__ int cidx(int k, int c, int C8) { return (k * C8 + (c >> 3)) * CPITCH + (c & 7); }

template <typename T, bool POOL, int ACT>
__global__ void __launch_bounds__(256, POOL ? 2 : 3) norm_act_fwd_kernel(const NormActFwdParams<T> p) {
  extern __shared__ float smem_f[];
  float* cst = smem_f;                     // [3][C]: scale, shift, prelu slope
  float* red = smem_f + 3 * (p.C >> 3) * CPITCH;   // [blockDim.x][8] when avg_sum
  const int n = blockIdx.y;
  const int C8 = p.C >> 3;
  const int H1 = p.H + 1;
  const int cd = POOL ? p.D / 2 : p.D, ch = POOL ? p.H / 2 : p.H, cw = POOL ? p.W / 2 : p.W;
  const int R = cd * (ch + 1);                 // merged rows of the (pooled) output
  const int cws = cw / p.wseg, Rv = R * p.wseg;   // segment width, virtual (row, segment) count
  const RowMap m = make_row_map(cws, C8);
  const bool lane_ok = m.t_row < m.rows_per_iter && m.chunks == 1;
  {
    const size_t o = p.per_sample ? (size_t)n * p.C : 0;
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      cst[cidx(K_SC, c, C8)] = p.scale[o + c];
      cst[cidx(K_SH, c, C8)] = p.shift[o + c];
      cst[cidx(K_SL, c, C8)] = p.prelu ? p.prelu[c] : 0.f;
    }
  }
  __syncthreads();
  const float* my = cst + m.c8 * CPITCH;
  const int KS = C8 * CPITCH;                  // floats between two constant arrays
  float asum[8];
#pragma unroll
  for (int i = 0; i < 8; i++) asum[i] = 0.f;
  constexpr int U = POOL ? 1 : 4;
  const size_t row_elems = (size_t)p.W * p.C;            // elements of one fine row
  const size_t base = (size_t)n * R * row_elems + m.c8 * 8;      // non-POOL: + row*row_elems + w*C
  // Row bookkeeping without integer division in the loop (rows no wider than a block, wseg == 1: every
  // 64x64x32 shape): row % period and row / period of the thread's first row are advanced incrementally;
  // the generic path divides (the divisions were ~2/3 of the instructions issued per row)
  const int rpi = m.rows_per_iter, stride = gridDim.x * rpi * U, period = POOL ? ch + 1 : H1;
  const bool w1 = p.wseg == 1;
  int rem0 = (blockIdx.x * rpi * U + m.t_row) % period, quo0 = (blockIdx.x * rpi * U + m.t_row) / period;
  const int d_rem = stride % period, d_quo = stride / period, u_rem = rpi % period;
  for (int row0 = blockIdx.x * rpi * U; row0 < Rv;
       row0 += stride, rem0 += d_rem, quo0 += d_quo + (rem0 >= period ? 1 : 0), rem0 -= (rem0 >= period ? period : 0)) {
    if (!POOL) {
      int rem = rem0;
      V8<T> yv[U];
      int kind[U];
      size_t off[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int vrow = row0 + u * rpi + m.t_row;
        int row, w, hp;
        if (w1) { row = vrow; w = m.w; hp = rem; rem += u_rem; rem -= (rem >= H1 ? H1 : 0); }
        else { row = vrow / p.wseg; w = (vrow % p.wseg) * cws + m.w; hp = row % H1; }
        kind[u] = 0;
        if (lane_ok && vrow < Rv) {
          off[u] = base + (size_t)row * row_elems + (size_t)w * p.C;
          kind[u] = hp == 0 ? 1 : 2;
          if (kind[u] == 2) yv[u] = ld8(p.y + off[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (kind[u] == 1) {
          if (p.a_out) z8(p.a_out + off[u]);
        } else if (kind[u] == 2) {
          float v[8], sc[8], sh[8], sl[8];
          up8(yv[u], v);
          ldc8(my + K_SC * KS, sc);
          ldc8(my + K_SH * KS, sh);
          if (ACT != ACT_RELU) ldc8(my + K_SL * KS, sl);
#pragma unroll
          for (int q = 0; q < 8; q++) v[q] = act_f<ACT>(fmaf(v[q], sc[q], sh[q]), p.act, ACT != ACT_RELU ? sl[q] : 0.f);
          rnd8(T(), v);  // average the stored (rounded) activations
          if (p.a_out) st8(p.a_out + off[u], v);
          if (p.avg_sum) {
#pragma unroll
            for (int q = 0; q < 8; q++) asum[q] += v[q];
          }
        }
      }
    } else {
      const int vrow = row0 + m.t_row;
      if (!(lane_ok && vrow < Rv)) continue;
      int row, w, hp, d;
      if (w1) { row = vrow; w = m.w; hp = rem0; d = quo0; }
      else { row = vrow / p.wseg; w = (vrow % p.wseg) * cws + m.w; hp = row % (ch + 1); d = row / (ch + 1); }
      const size_t pooled_off = (((size_t)n * R + row) * cw + w) * p.C + m.c8 * 8;
      if (hp == 0) {  // pad rows of the outputs
        if (p.pool_out) z8(p.pool_out + pooled_off);
        if (p.a_out)
          for (int i = 0; i < 2; i++)
            for (int k = 0; k < 2; k++)
              z8(p.a_out + ((((size_t)n * p.D + 2 * d + i) * H1) * p.W + 2 * w + k) * p.C + m.c8 * 8);
        continue;
      }
      const int h = hp - 1;
      V8<T> yv[8];
#pragma unroll
      for (int pos = 0; pos < 8; pos++) {
        const int i = pos >> 2, j = (pos >> 1) & 1, k = pos & 1;
        yv[pos] = ld8(
            p.y + ((((size_t)n * p.D + 2 * d + i) * H1 + 2 * h + j + 1) * p.W + 2 * w + k) * p.C + m.c8 * 8);
      }
      float mx[8], sc[8], sh[8], sl[8];
      ldc8(my + K_SC * KS, sc);
      ldc8(my + K_SH * KS, sh);
      if (ACT != ACT_RELU) ldc8(my + K_SL * KS, sl);
#pragma unroll
      for (int i = 0; i < 8; i++) mx[i] = -INFINITY;
#pragma unroll
      for (int pos = 0; pos < 8; pos++) {
        const int i = pos >> 2, j = (pos >> 1) & 1, k = pos & 1;
        float v[8];
        up8(yv[pos], v);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          v[q] = act_f<ACT>(fmaf(v[q], sc[q], sh[q]), p.act, ACT != ACT_RELU ? sl[q] : 0.f);
          mx[q] = fmaxf(mx[q], v[q]);
        }
        if (p.a_out)
          st8(p.a_out + ((((size_t)n * p.D + 2 * d + i) * H1 + 2 * h + j + 1) * p.W + 2 * w + k) * p.C + m.c8 * 8, v);
      }
      if (p.pool_out) st8(p.pool_out + pooled_off, mx);
    }
  }
  if (p.avg_sum) {
#pragma unroll
    for (int q = 0; q < 8; q++) red[threadIdx.x * 8 + q] = lane_ok ? asum[q] : 0.f;
    __syncthreads();
    if (threadIdx.x < C8) {
      // threads with the same channel group: t_item % C8 == c8, i.e. threadIdx % C8 when
      // items_per_row is a multiple of C8 (it is: items_per_row = cw * C8)
      float t[8];
#pragma unroll
      for (int q = 0; q < 8; q++) t[q] = 0.f;
      for (int j = threadIdx.x; j < blockDim.x; j += C8)
#pragma unroll
        for (int q = 0; q < 8; q++) t[q] += red[j * 8 + q];
#pragma unroll
      for (int q = 0; q < 8; q++) atomicAdd(&p.avg_sum[(size_t)n * p.C + threadIdx.x * 8 + q], t[q]);
    }
  }
}

// ------------------------------------------------------------------------------ norm + act backward
// Upstream gradient wrt a = act(z), z = y*scale + shift:
//   g1: H-padded, full resolution, or (POOL) the gradient wrt the 2x2x2 max-pooled tensor,
//       routed to the first maximum of each cell (torch's MaxPool3d tie rule);
//   g2: optional second full-resolution gradient, added (G2: compile-time, only the tail layers have it);
//   gavg: optional [N][C] fp32 gradient wrt the global average pool; adds gavg/(D*H*W).
// pass 1 accumulates per (group, channel): sum dz, sum dz*xhat (and sum dA*min(z,0) for PReLU);
// pass 2 writes dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)), pad rows zero.
template <typename T>
struct NormActBwdParams {
  const T* y; const T* g1; const T* g2; const float* gavg;
  const float* scale; const float* shift; const float* mean; const float* invstd;
  const float* gamma; const float* prelu;
  double* sums;      // [G][C][3]
  T* dy; // pass 2
  double count;      // elements per (group, channel)
  int per_sample, act, N, D, H, W, C;
  int wseg;          // a row of W voxels is processed as wseg segments (rows wider than one block)
};

template <typename T, bool POOL, bool APPLY, int ACT, bool G2>
__global__ void __launch_bounds__(256, POOL ? 2 : 3) norm_act_bwd_kernel(const NormActBwdParams<T> p) {
  extern __shared__ float smem_f[];
  float* cst = smem_f;                       // [K_COUNT][C]
  float* red = smem_f + K_COUNT * (p.C >> 3) * CPITCH;       // pass 1: [blockDim.x][24]
  const int n = blockIdx.y;
  const int C = p.C, C8 = p.C >> 3;
  const int H1 = p.H + 1;
  const int cd = POOL ? p.D / 2 : p.D, ch = POOL ? p.H / 2 : p.H, cw = POOL ? p.W / 2 : p.W;
  const int R = cd * (ch + 1);
  const int cws = cw / p.wseg, Rv = R * p.wseg;   // segment width, virtual (row, segment) count
  const RowMap m = make_row_map(cws, C8);
  const bool lane_ok = m.t_row < m.rows_per_iter && m.chunks == 1;
  // x_hat = y*is - mis ; z = y*sc + sh ; apply: dy = gs*dz - c1 - y*c2   (constants folded, one value per
  // channel computed by one thread)
  {
    const size_t o = p.per_sample ? (size_t)n * C : 0;
    const float inv_vol = 1.f / ((float)p.D * p.H * p.W);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float is = p.invstd[o + c], mis = p.mean[o + c] * is;
      cst[cidx(K_SC, c, C8)] = p.scale[o + c];
      cst[cidx(K_SH, c, C8)] = p.shift[o + c];
      cst[cidx(K_SL, c, C8)] = p.prelu ? p.prelu[c] : 0.f;
      if (!APPLY) {
        cst[cidx(K_IS, c, C8)] = is;
        cst[cidx(K_MIS, c, C8)] = mis;
      }
      cst[cidx(K_GA, c, C8)] = p.gavg ? p.gavg[(size_t)n * C + c] * inv_vol : 0.f;
      if (APPLY) {
        const double* sm = p.sums + (o + c) * 3;
        const float k1 = (float)(sm[0] / p.count), k2 = (float)(sm[1] / p.count);
        const float gs = p.gamma[c] * is;
        cst[cidx(K_GS, c, C8)] = gs;
        cst[cidx(K_C1, c, C8)] = gs * (k1 - mis * k2);      // gs*(dz - k1 - xh*k2) with xh = y*is - mis
        cst[cidx(K_C2, c, C8)] = gs * k2 * is;
      }
    }
  }
  __syncthreads();
  const float* my = cst + m.c8 * CPITCH;
  const int KS = C8 * CPITCH;                  // floats between two constant arrays
  const bool has_ga = p.gavg != nullptr;
  float a0[8], a1[8], a2[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a0[i] = a1[i] = a2[i] = 0.f;
  constexpr int U = POOL ? 1 : (sizeof(T) == 2 ? 4 : 2);     // see norm_act_fwd_kernel
  const size_t row_elems = (size_t)p.W * C;
  const size_t base = (size_t)n * R * row_elems + m.c8 * 8;      // non-POOL: + row*row_elems + w*C
  // Row bookkeeping without integer division in the loop (rows no wider than a block, wseg == 1: every
  // 64x64x32 shape): row % period and row / period of the thread's first row are advanced incrementally;
  // the generic path divides (the divisions were ~2/3 of the instructions issued per row)
  const int rpi = m.rows_per_iter, stride = gridDim.x * rpi * U, period = POOL ? ch + 1 : H1;
  const bool w1 = p.wseg == 1;
  int rem0 = (blockIdx.x * rpi * U + m.t_row) % period, quo0 = (blockIdx.x * rpi * U + m.t_row) / period;
  const int d_rem = stride % period, d_quo = stride / period, u_rem = rpi % period;
  for (int row0 = blockIdx.x * rpi * U; row0 < Rv;
       row0 += stride, rem0 += d_rem, quo0 += d_quo + (rem0 >= period ? 1 : 0), rem0 -= (rem0 >= period ? period : 0)) {
    if (!POOL) {
      int rem = rem0;
      V8<T> yq[U], g1q[U], g2q[U];
      int kind[U];
      size_t off[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int vrow = row0 + u * rpi + m.t_row;
        int row, w, hp;
        if (w1) { row = vrow; w = m.w; hp = rem; rem += u_rem; rem -= (rem >= H1 ? H1 : 0); }
        else { row = vrow / p.wseg; w = (vrow % p.wseg) * cws + m.w; hp = row % H1; }
        kind[u] = 0;
        if (lane_ok && vrow < Rv) {
          off[u] = base + (size_t)row * row_elems + (size_t)w * C;
          kind[u] = hp == 0 ? 1 : 2;
          if (kind[u] == 2) {
            yq[u] = ld8(p.y + off[u]);
            if (p.g1) g1q[u] = ld8(p.g1 + off[u]);
            if (G2) g2q[u] = ld8(p.g2 + off[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (kind[u] == 1) {
          if (APPLY) z8(p.dy + off[u]);
        } else if (kind[u] == 2) {
          float yv[8], da[8], t[8], z[8];
          up8(yq[u], yv);
          if (has_ga) ldc8(my + K_GA * KS, da);
          else {
#pragma unroll
            for (int q = 0; q < 8; q++) da[q] = 0.f;
          }
          if (p.g1) {
            up8(g1q[u], t);
#pragma unroll
            for (int q = 0; q < 8; q++) da[q] += t[q];
          }
          if (G2) {
            up8(g2q[u], t);
#pragma unroll
            for (int q = 0; q < 8; q++) da[q] += t[q];
          }
          {
            float sc[8], sh[8];
            ldc8(my + K_SC * KS, sc);
            ldc8(my + K_SH * KS, sh);
#pragma unroll
            for (int q = 0; q < 8; q++) z[q] = fmaf(yv[q], sc[q], sh[q]);
          }
          if (ACT != ACT_RELU) {
            ldc8(my + K_SL * KS, t);
            if (!APPLY) {
#pragma unroll
              for (int q = 0; q < 8; q++) a2[q] += da[q] * fminf(z[q], 0.f);
            }
#pragma unroll
            for (int q = 0; q < 8; q++) da[q] *= act_bwd(z[q], p.act, t[q]);      // da now holds dz
          } else {
#pragma unroll
            for (int q = 0; q < 8; q++) da[q] = z[q] > 0.f ? da[q] : 0.f;
          }
          if (APPLY) {
            float gs[8], c1[8], c2[8];
            ldc8(my + K_GS * KS, gs);
            ldc8(my + K_C1 * KS, c1);
            ldc8(my + K_C2 * KS, c2);
#pragma unroll
            for (int q = 0; q < 8; q++) z[q] = fmaf(gs[q], da[q], -fmaf(yv[q], c2[q], c1[q]));
            st8(p.dy + off[u], z);
          } else {
            float is[8], mis[8];
            ldc8(my + K_IS * KS, is);
            ldc8(my + K_MIS * KS, mis);
#pragma unroll
            for (int q = 0; q < 8; q++) {
              a0[q] += da[q];
              a1[q] = fmaf(da[q], fmaf(yv[q], is[q], -mis[q]), a1[q]);
            }
          }
        }
      }
    } else {
      const int vrow = row0 + m.t_row;
      if (!(lane_ok && vrow < Rv)) continue;
      int row, w, hp, d;
      if (w1) { row = vrow; w = m.w; hp = rem0; d = quo0; }
      else { row = vrow / p.wseg; w = (vrow % p.wseg) * cws + m.w; hp = row % (ch + 1); d = row / (ch + 1); }
      if (hp == 0) {
        if (APPLY)
          for (int i = 0; i < 2; i++)
            for (int k = 0; k < 2; k++)
              z8(p.dy + ((((size_t)n * p.D + 2 * d + i) * H1) * p.W + 2 * w + k) * C + m.c8 * 8);
        continue;
      }
      const int h = hp - 1;
      float gp[8];
      up8(ld8(p.g1 + (((size_t)n * R + row) * cw + w) * C + m.c8 * 8), gp);
      V8<T> yq[8];
#pragma unroll
      for (int pos = 0; pos < 8; pos++) {
        const int i = pos >> 2, j = (pos >> 1) & 1, k = pos & 1;
        yq[pos] = ld8(
            p.y + ((((size_t)n * p.D + 2 * d + i) * H1 + 2 * h + j + 1) * p.W + 2 * w + k) * C + m.c8 * 8);
      }
      float sc[8], sh[8], sl[8], ga[8];
      ldc8(my + K_SC * KS, sc);
      ldc8(my + K_SH * KS, sh);
      if (ACT != ACT_RELU) ldc8(my + K_SL * KS, sl);
      if (has_ga) ldc8(my + K_GA * KS, ga);
      int arg[8];
      float mx[8];
#pragma unroll
      for (int q = 0; q < 8; q++) { mx[q] = -INFINITY; arg[q] = 0; }
#pragma unroll
      for (int pos = 0; pos < 8; pos++) {
        float yv[8];
        up8(yq[pos], yv);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          // same fp32 values and scan order as the forward pass: the first maximum wins
          const float a = act_f<ACT>(fmaf(yv[q], sc[q], sh[q]), p.act, ACT != ACT_RELU ? sl[q] : 0.f);
          if (a > mx[q]) { mx[q] = a; arg[q] = pos; }
        }
      }
#pragma unroll
      for (int pos = 0; pos < 8; pos++) {
        const int i = pos >> 2, j = (pos >> 1) & 1, k = pos & 1;
        const size_t off = ((((size_t)n * p.D + 2 * d + i) * H1 + 2 * h + j + 1) * p.W + 2 * w + k) * C + m.c8 * 8;
        float yv[8], g2v[8], out[8], dzv[8];
        // y is re-read here (an L1 / L2 hit: this thread loaded the same 8 vectors a moment ago) instead of
        // keeping all eight positions live across both loops, which spilled at 2 blocks per SM
        up8(ld8(p.y + off), yv);
        if (G2) up8(ld8(p.g2 + off), g2v);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float z = fmaf(yv[q], sc[q], sh[q]);
          float da = (arg[q] == pos ? gp[q] : 0.f) + (has_ga ? ga[q] : 0.f);
          if (G2) da += g2v[q];
          dzv[q] = da * act_d<ACT>(z, p.act, ACT != ACT_RELU ? sl[q] : 0.f);
          if (!APPLY && ACT != ACT_RELU) a2[q] += da * fminf(z, 0.f);
        }
        if (APPLY) {
          float gs[8], c1[8], c2[8];
          ldc8(my + K_GS * KS, gs);
          ldc8(my + K_C1 * KS, c1);
          ldc8(my + K_C2 * KS, c2);
#pragma unroll
          for (int q = 0; q < 8; q++) out[q] = fmaf(gs[q], dzv[q], -fmaf(yv[q], c2[q], c1[q]));
          st8(p.dy + off, out);
        } else {
          float is[8], mis[8];
          ldc8(my + K_IS * KS, is);
          ldc8(my + K_MIS * KS, mis);
#pragma unroll
          for (int q = 0; q < 8; q++) {
            a0[q] += dzv[q];
            a1[q] = fmaf(dzv[q], fmaf(yv[q], is[q], -mis[q]), a1[q]);
          }
        }
      }
    }
  }
  if (!APPLY) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      red[threadIdx.x * 24 + q] = lane_ok ? a0[q] : 0.f;
      red[threadIdx.x * 24 + 8 + q] = lane_ok ? a1[q] : 0.f;
      red[threadIdx.x * 24 + 16 + q] = lane_ok ? a2[q] : 0.f;
    }
    __syncthreads();
    if (threadIdx.x < C8 * 3) {
      const int grp = threadIdx.x % C8, which = threadIdx.x / C8;
      float t[8];
#pragma unroll
      for (int q = 0; q < 8; q++) t[q] = 0.f;
      for (int j = grp; j < blockDim.x; j += C8)
#pragma unroll
        for (int q = 0; q < 8; q++) t[q] += red[j * 24 + which * 8 + q];
      const size_t o = (p.per_sample ? (size_t)n * C : 0) + grp * 8;
#pragma unroll
      for (int q = 0; q < 8; q++) atomicAdd(&p.sums[(o + q) * 3 + which], (double)t[q]);
    }
  }
}

// ------------------------------------------------------------------------------ misc
// zero the pad row (h' = 0) of every plane of an H-padded tensor; row_vecs = 16-byte vectors per row
__global__ void zero_pad_rows_kernel(uint4* __restrict__ t, long long planes, int H1, int row_vecs) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < planes * row_vecs;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pl = i / row_vecs;
    const int e = (int)(i % row_vecs);
    t[(size_t)pl * H1 * row_vecs + e] = make_uint4(0, 0, 0, 0);
  }
}

// ConvTranspose backward re-layout: fine gradient [N][2D][2H+1][2W][C] bf16 -> coarse-major
// [N*D*(H+1)*W][8*C] bf16 (rows follow the coarse H-padded order, pad rows zero), plus the bias
// gradient dbias[C] += sum over all fine voxels.
template <typename T>
__global__ void __launch_bounds__(256)
convT_unshuffle_kernel(const T* __restrict__ g, T* __restrict__ out,
                       float* __restrict__ dbias, int N, int D, int H, int W, int C) {
  extern __shared__ float red[];  // [blockDim.x][8]
  const int C8 = C >> 3;
  const long long rows = (long long)N * D * (H + 1) * W;
  const long long items = rows * 8 * C8;
  const int c8 = threadIdx.x % C8;
  float bs[8];
#pragma unroll
  for (int i = 0; i < 8; i++) bs[i] = 0.f;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int t = (int)((it / C8) % 8);
    const long long r = it / (8 * C8);
    const int w = (int)(r % W);
    const int hp = (int)((r / W) % (H + 1));
    const int d = (int)((r / ((long long)W * (H + 1))) % D);
    const long long n = r / ((long long)W * (H + 1) * D);
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; q++) f[q] = 0.f;
    if (hp >= 1) {
      const int i = t >> 2, j = (t >> 1) & 1, k = t & 1;
      const size_t off = ((((size_t)n * 2 * D + 2 * d + i) * (2 * H + 1) + 2 * (hp - 1) + j + 1) * (2 * W) + 2 * w + k) * C + c8 * 8;
      up8(ld8(g + off), f);
#pragma unroll
      for (int q = 0; q < 8; q++) bs[q] += f[q];
    }
    st8(out + ((size_t)r * 8 + t) * C + c8 * 8, f);
  }
  if (dbias) {
#pragma unroll
    for (int q = 0; q < 8; q++) red[threadIdx.x * 8 + q] = bs[q];
    __syncthreads();
    if (threadIdx.x < C8) {
      float t[8];
#pragma unroll
      for (int q = 0; q < 8; q++) t[q] = 0.f;
      for (int j = threadIdx.x; j < blockDim.x; j += C8)
#pragma unroll
        for (int q = 0; q < 8; q++) t[q] += red[j * 8 + q];
#pragma unroll
      for (int q = 0; q < 8; q++) atomicAdd(&dbias[threadIdx.x * 8 + q], t[q]);
    }
  }
}

// ------------------------------------------------------------------------------ 3xTF32 operand split
// src [rows][C] fp32 -> (hi, lo, hi) or (hi, hi, lo) with hi = rna_tf32(x), lo = rna_tf32(x - hi), written
// as [rows][3C] (parts along the contraction index) or [3][rows][C] (parts along the row index).
__global__ void __launch_bounds__(256)
split3_tf32_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long rows, int C4,
                   int pattern, int stack_rows) {
  const long long total = rows * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C4;
    const int c = (int)(i % C4);
    const float4 x = src[i];
    float4 hi, lo;
    hi.x = rna_tf32(x.x); hi.y = rna_tf32(x.y); hi.z = rna_tf32(x.z); hi.w = rna_tf32(x.w);
    lo.x = rna_tf32(x.x - hi.x); lo.y = rna_tf32(x.y - hi.y); lo.z = rna_tf32(x.z - hi.z); lo.w = rna_tf32(x.w - hi.w);
    const float4 p1 = pattern == 0 ? lo : hi, p2 = pattern == 0 ? hi : lo;
    if (stack_rows) {
      dst[i] = hi;
      dst[total + i] = p1;
      dst[2 * total + i] = p2;
    } else {
      float4* o = dst + r * 3 * C4 + c;
      o[0] = hi;
      o[C4] = p1;
      o[2 * C4] = p2;
    }
  }
}

// ------------------------------------------------------------------------------ launch wrappers
static inline int grid_for(long long items, int block, int max_blocks) {
  long long b = (items + block - 1) / block;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (int)b;
}
// block size: a multiple of C/8 near 256 so that a thread's channel group is loop-invariant
static inline int block_for_c8(int C8) {
  int b = (256 / C8) * C8;
  if (b == 0) b = C8;  // C8 > 256 cannot happen for C <= 2048
  return b;
}

// dtype: PCRL_DTYPE_BF16 (0) or PCRL_DTYPE_F32 (1) = storage type of activations / packed operands
#define PCRL_BY_DTYPE(dtype, EXPR_BF16, EXPR_F32, EXPR_F32X) \
  do { if ((dtype) == PCRL_DTYPE_F32) { EXPR_F32; } else if ((dtype) == PCRL_DTYPE_F32X) { EXPR_F32X; } else { EXPR_BF16; } } while (0)

int split3_tf32(const float* src, float* dst, long long rows, int C, int pattern, int stack_rows, cudaStream_t s) {
  PCRL_REQUIRE(C % 4 == 0 && rows > 0, "split3_tf32: C=%d must be a multiple of 4", C);
  PCRL_REQUIRE(pattern == 0 || pattern == 1, "split3_tf32: pattern must be 0 (hi,lo,hi) or 1 (hi,hi,lo)");
  split3_tf32_kernel<<<grid_for(rows * (C / 4), 256, num_sms() * 16), 256, 0, s>>>(
      (const float4*)src, (float4*)dst, rows, C / 4, pattern, stack_rows);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int pack_conv3_weights(const float* w, void* wf, void* wd, int Cout, int Cin, int dtype, cudaStream_t s) {
  const long long total = (long long)Cout * Cin * 27;
  PCRL_BY_DTYPE(dtype,
    (pack_conv3_kernel<bf16_t><<<grid_for(total, 256, 4096), 256, 0, s>>>(w, (bf16_t*)wf, (bf16_t*)wd, Cout, Cin)),
    (pack_conv3_kernel<float><<<grid_for(total, 256, 4096), 256, 0, s>>>(w, (float*)wf, (float*)wd, Cout, Cin)),
    (pack_conv3_kernel<f32x><<<grid_for(total, 256, 4096), 256, 0, s>>>(w, (f32x*)wf, (f32x*)wd, Cout, Cin)));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int unpack_conv3_wgrad(const float* gpk, float* g, int Cout, int Cin, cudaStream_t s) {
  const long long total = (long long)Cout * Cin * 27;
  unpack_conv3_wgrad_kernel<<<grid_for(total, 256, 4096), 256, 0, s>>>(gpk, g, Cout, Cin);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int pack_convT_weights(const float* w, void* wf, void* wd, int Cin, int Cout, int dtype, cudaStream_t s) {
  const long long total = (long long)Cin * Cout * 8;
  PCRL_BY_DTYPE(dtype,
    (pack_convT_kernel<bf16_t><<<grid_for(total, 256, 4096), 256, 0, s>>>(w, (bf16_t*)wf, (bf16_t*)wd, Cin, Cout)),
    (pack_convT_kernel<float><<<grid_for(total, 256, 4096), 256, 0, s>>>(w, (float*)wf, (float*)wd, Cin, Cout)),
    (pack_convT_kernel<f32x><<<grid_for(total, 256, 4096), 256, 0, s>>>(w, (f32x*)wf, (f32x*)wd, Cin, Cout)));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int unpack_convT_wgrad(const float* gpk, float* g, int Cin, int Cout, cudaStream_t s) {
  const long long total = (long long)Cin * Cout * 8;
  unpack_convT_wgrad_kernel<<<grid_for(total, 256, 4096), 256, 0, s>>>(gpk, g, Cin, Cout);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int stem_conv_fprop(const float* x, const float* w, void* y, double* stats, int stats_per_sample,
                    int N, int D, int H, int W, int dtype, cudaStream_t s) {
  const int slots = D * (H + 1) * W;
  dim3 grid((slots + 127) / 128, N);
  PCRL_BY_DTYPE(dtype,
    (stem_conv_fprop_kernel<bf16_t><<<grid, 128, 0, s>>>(x, w, (bf16_t*)y, stats, stats_per_sample, N, D, H, W)),
    (stem_conv_fprop_kernel<float><<<grid, 128, 0, s>>>(x, w, (float*)y, stats, stats_per_sample, N, D, H, W)),
    (stem_conv_fprop_kernel<f32x><<<grid, 128, 0, s>>>(x, w, (f32x*)y, stats, stats_per_sample, N, D, H, W)));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}
int stem_conv_wgrad(const void* dy, const float* x, float* dw, int N, int D, int H, int W, int dtype,
                    cudaStream_t s) {
  const long long total = (long long)N * D * H * W;
  const int warps = num_sms() * 16;
  int vpw = (int)((total + warps - 1) / warps);
  if (vpw < 1) vpw = 1;
  const long long nwarps = (total + vpw - 1) / vpw;
  const int blocks = (int)((nwarps + 7) / 8);
  PCRL_BY_DTYPE(dtype,
    (stem_conv_wgrad_kernel<bf16_t><<<blocks, 256, 0, s>>>((const bf16_t*)dy, x, dw, N, D, H, W, vpw)),
    (stem_conv_wgrad_kernel<float><<<blocks, 256, 0, s>>>((const float*)dy, x, dw, N, D, H, W, vpw)),
    (stem_conv_wgrad_kernel<f32x><<<blocks, 256, 0, s>>>((const f32x*)dy, x, dw, N, D, H, W, vpw)));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int norm_finalize(const double* stats, double count, const float* gamma, const float* beta,
                  const float* conv_bias, float* running_mean, float* running_var,
                  long long* nbt, float momentum, float eps, float* scale, float* shift,
                  float* mean, float* invstd, int G, int C, cudaStream_t s) {
  norm_finalize_kernel<<<(G * C + 127) / 128, 128, 0, s>>>(stats, count, gamma, beta, conv_bias, running_mean,
                                                          running_var, nbt, momentum, eps, scale, shift,
                                                          mean, invstd, G, C);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// number of equal segments a row of cw voxels x C8 8-channel groups is cut into so that one segment
// fits a 256-thread block (1 for every 64x64x32 layer; 2..4 for the 128x128x64 crops); 0 = impossible
static int row_segments(int cw, int C8) {
  for (int sgm = 1; sgm <= cw; sgm++)
    if (cw % sgm == 0 && (cw / sgm) * C8 <= 256) return sgm;
  return 0;
}

template <typename T>
static int norm_act_fwd_t(const void* y, const float* scale, const float* shift, const float* prelu,
                          void* a_out, void* pool_out, float* avg_sum, int per_sample, int act, int pool,
                          int N, int D, int H, int W, int C, cudaStream_t s) {
  const int cw = pool ? W / 2 : W;
  const int wseg = row_segments(cw, C / 8);
  NormActFwdParams<T> p{(const T*)y, scale, shift, prelu, (T*)a_out, (T*)pool_out, avg_sum,
                        per_sample, act, N, D, H, W, C, wseg};
  const int items = (cw / wseg) * (C / 8), rpi = 256 / items, U = pool ? 1 : 4;
  const int R = (pool ? (D / 2) * (H / 2 + 1) : D * (H + 1)) * wseg;
  int bx = (R + rpi * U - 1) / (rpi * U);
  const int cap = (num_sms() * 8 + N - 1) / N;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, N);
  const size_t smem = (size_t)3 * (C / 8) * CPITCH * 4 + (avg_sum ? (size_t)256 * 8 * 4 : 0);
  const bool relu = act == ACT_RELU;
  if (pool) {
    if (relu) norm_act_fwd_kernel<T, true, ACT_RELU><<<grid, 256, smem, s>>>(p);
    else norm_act_fwd_kernel<T, true, -1><<<grid, 256, smem, s>>>(p);
  } else {
    if (relu) norm_act_fwd_kernel<T, false, ACT_RELU><<<grid, 256, smem, s>>>(p);
    else norm_act_fwd_kernel<T, false, -1><<<grid, 256, smem, s>>>(p);
  }
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int norm_act_fwd(const void* y, const float* scale, const float* shift, const float* prelu,
                 void* a_out, void* pool_out, float* avg_sum, int per_sample, int act, int pool,
                 int N, int D, int H, int W, int C, int dtype, cudaStream_t s) {
  PCRL_REQUIRE(C % 8 == 0 && ((C / 8) & (C / 8 - 1)) == 0, "norm_act_fwd: C=%d must be 8 * 2^k", C);
  PCRL_REQUIRE(!pool || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "norm_act_fwd: pooling needs even dims");
  PCRL_REQUIRE(!(pool && avg_sum), "norm_act_fwd: avg_sum with pool is not supported");
  PCRL_REQUIRE(row_segments(pool ? W / 2 : W, C / 8) > 0, "norm_act_fwd: row of %d x %d channels cannot be split into block-sized segments", W, C);
  if (dtype == PCRL_DTYPE_F32X)
    return norm_act_fwd_t<f32x>(y, scale, shift, prelu, a_out, pool_out, avg_sum, per_sample, act, pool, N, D, H, W, C, s);
  if (dtype == PCRL_DTYPE_F32)
    return norm_act_fwd_t<float>(y, scale, shift, prelu, a_out, pool_out, avg_sum, per_sample, act, pool, N, D, H, W, C, s);
  return norm_act_fwd_t<bf16_t>(y, scale, shift, prelu, a_out, pool_out, avg_sum, per_sample, act, pool, N, D, H, W, C, s);
}

template <typename T>
static int norm_act_bwd_t(const void* y, const void* g1, const void* g2, const float* gavg,
                          const float* scale, const float* shift, const float* mean, const float* invstd,
                          const float* gamma, const float* prelu, double* sums, void* dy, double count,
                          int per_sample, int act, int pool, int pass, int N, int D, int H, int W, int C,
                          cudaStream_t s) {
  const int cw = pool ? W / 2 : W;
  NormActBwdParams<T> p{(const T*)y, (const T*)g1, (const T*)g2, gavg, scale, shift, mean, invstd, gamma,
                        prelu, sums, (T*)dy, count, per_sample, act, N, D, H, W, C, row_segments(cw, C / 8)};
  const int items = (cw / p.wseg) * (C / 8), rpi = 256 / items, U = pool ? 1 : (sizeof(T) == 2 ? 4 : 2);
  const int R = (pool ? (D / 2) * (H / 2 + 1) : D * (H + 1)) * p.wseg;
  int bx = (R + rpi * U - 1) / (rpi * U);
  const int cap = (num_sms() * 8 + N - 1) / N;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, N);
  const bool relu = act == ACT_RELU;
  const bool has_g2 = g2 != nullptr;
  const size_t csm = (size_t)K_COUNT * (C / 8) * CPITCH * 4;
#define PCRL_LAUNCH_BWD2(POOLV, APPLYV, SMEM, G2V)                                                \
  do {                                                                                            \
    if (relu) norm_act_bwd_kernel<T, POOLV, APPLYV, ACT_RELU, G2V><<<grid, 256, SMEM, s>>>(p);    \
    else norm_act_bwd_kernel<T, POOLV, APPLYV, -1, G2V><<<grid, 256, SMEM, s>>>(p);               \
  } while (0)
#define PCRL_LAUNCH_BWD(POOLV, APPLYV, SMEM)                    \
  do {                                                          \
    if (has_g2) PCRL_LAUNCH_BWD2(POOLV, APPLYV, SMEM, true);    \
    else PCRL_LAUNCH_BWD2(POOLV, APPLYV, SMEM, false);          \
  } while (0)
  if (pass == 0) {
    const size_t smem = csm + (size_t)256 * 24 * 4;
    if (pool) PCRL_LAUNCH_BWD(true, false, smem);
    else PCRL_LAUNCH_BWD(false, false, smem);
  } else {
    if (pool) PCRL_LAUNCH_BWD(true, true, csm);
    else PCRL_LAUNCH_BWD(false, true, csm);
  }
#undef PCRL_LAUNCH_BWD2
#undef PCRL_LAUNCH_BWD
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// pass = 0: reduce into sums [G][C][3] (caller zeroes), pass = 1: apply (writes dy)
int norm_act_bwd(const void* y, const void* g1, const void* g2, const float* gavg,
                 const float* scale, const float* shift, const float* mean, const float* invstd,
                 const float* gamma, const float* prelu, double* sums, void* dy, double count,
                 int per_sample, int act, int pool, int pass, int N, int D, int H, int W, int C,
                 int dtype, cudaStream_t s) {
  PCRL_REQUIRE(C % 8 == 0 && ((C / 8) & (C / 8 - 1)) == 0, "norm_act_bwd: C=%d must be 8 * 2^k", C);
  PCRL_REQUIRE(!pool || g1, "norm_act_bwd: pooled backward needs g1");
  PCRL_REQUIRE(row_segments(pool ? W / 2 : W, C / 8) > 0, "norm_act_bwd: row of %d x %d channels cannot be split into block-sized segments", W, C);
  if (dtype == PCRL_DTYPE_F32X)
    return norm_act_bwd_t<f32x>(y, g1, g2, gavg, scale, shift, mean, invstd, gamma, prelu, sums, dy, count,
                                per_sample, act, pool, pass, N, D, H, W, C, s);
  if (dtype == PCRL_DTYPE_F32)
    return norm_act_bwd_t<float>(y, g1, g2, gavg, scale, shift, mean, invstd, gamma, prelu, sums, dy, count,
                                 per_sample, act, pool, pass, N, D, H, W, C, s);
  return norm_act_bwd_t<bf16_t>(y, g1, g2, gavg, scale, shift, mean, invstd, gamma, prelu, sums, dy, count,
                                per_sample, act, pool, pass, N, D, H, W, C, s);
}

// row_bytes = bytes of one (w, c) row of the H-padded tensor
int zero_pad_rows(void* t, long long planes, int H1, long long row_bytes, cudaStream_t s) {
  PCRL_REQUIRE(row_bytes % 16 == 0, "zero_pad_rows: row of %lld bytes must be a multiple of 16", row_bytes);
  const int row_vecs = (int)(row_bytes / 16);
  zero_pad_rows_kernel<<<grid_for(planes * row_vecs, 256, 8192), 256, 0, s>>>((uint4*)t, planes, H1, row_vecs);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

int convT_unshuffle(const void* g, void* out, float* dbias, int N, int D, int H, int W, int C, int dtype,
                    cudaStream_t s) {
  PCRL_REQUIRE(C % 8 == 0, "convT_unshuffle: C=%d must be a multiple of 8", C);
  const int C8 = C / 8, block = block_for_c8(C8);
  const long long items = (long long)N * D * (H + 1) * W * 8 * C8;
  const int blocks = grid_for((items + 3) / 4, block, 1 << 20);
  const size_t smem = dbias ? (size_t)block * 8 * 4 : 0;
  PCRL_BY_DTYPE(dtype,
    (convT_unshuffle_kernel<bf16_t><<<blocks, block, smem, s>>>((const bf16_t*)g, (bf16_t*)out, dbias, N, D, H, W, C)),
    (convT_unshuffle_kernel<float><<<blocks, block, smem, s>>>((const float*)g, (float*)out, dbias, N, D, H, W, C)),
    (convT_unshuffle_kernel<f32x><<<blocks, block, smem, s>>>((const f32x*)g, (f32x*)out, dbias, N, D, H, W, C)));
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

}  // namespace pcrl
