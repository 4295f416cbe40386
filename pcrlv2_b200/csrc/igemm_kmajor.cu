// K-major implicit-GEMM on tcgen05 (sm_100a):   OUT[row][col] = sum_tap sum_k A_tap[row][k] * B[tap][col][k]
//
// One kernel serves
//   * 3x3x3 convolution forward  (A = activations, B = W packed [27][Cout][Cin])          -- LUConv,
//     reference models/pcrlv2_model_3d.py:9,33 (nn.Conv3d(k=3, padding=1))
//   * its data gradient          (A = dY,          B = W packed [27][Cin][Cout], taps flipped)
//   * ConvTranspose3d(k=2,s=2) forward as a GEMM with a scatter epilogue (reference :52,64)
//   * plain row-major GEMMs (ConvTranspose data gradient, Linear layers).
//
// CONV mode works on the per-sample flat index space of the H-padded NDHWC layout (common.cuh):
// a CTA owns m_cta consecutive flat rows; for every (k-block, dz) it TMA-loads ONE slab of whole
// merged rows (box (kc, W+1, nh, 1), zero halo by out-of-bounds fill) and issues the nine (dy,dx)
// taps as row-shifted views of that slab (descriptor start address + shift*row_bytes; verified on
// hardware, profiles/r01_umma_probe.md).  Weights stream through a separate TMA ring, one
// [nc x kc] tile per tap.  Accumulators live in TMEM (mt tiles of 128 rows x nc columns).
// Epilogue: tcgen05.ld -> (bias) -> bf16/fp32 store, plus per-channel sum / sum-of-squares for
// BatchNorm / InstanceNorm statistics (warp-shuffle transpose-reduce, fp64 global atomics).
#include "common.cuh"
#include "sm100.cuh"

namespace pcrl {

enum { IG_CONV = 0, IG_PLAIN = 1 };
enum { OUT_FLAT = 0, OUT_ROWS = 1, OUT_CONVT = 2 };

struct IgemmParams {
  int mode;
  // geometry (CONV)
  int W, Wp, H1, D, MR;   // H1 = H+1, MR = D*H1 merged rows per sample
  int nh_box;
  // tiling
  int m_cta, mt, nc, kc, row_bytes, kblocks, groups, tpg;
  int slab_bytes, b_bytes, sa, sb, tmem_cols;
  int b_rows_per_tap;
  long long rows_total;   // PLAIN: number of A rows
  // epilogue
  int out_mode, out_fp32, ldc, has_bias, has_stats, stats_per_sample, cout_total;
  int ct_D, ct_H, ct_W;   // coarse dims for the ConvT scatter
  void* out;
  const float* bias;
  double* stats;
};

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

__global__ void __launch_bounds__(128)
igemm_kmajor_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                    const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;
  uint8_t* b_s = a_s + (size_t)p.sa * p.slab_bytes;
  uint64_t* bars = (uint64_t*)(b_s + (size_t)p.sb * p.b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.sa;
  uint64_t* b_full = a_empty + p.sa;
  uint64_t* b_empty = b_full + p.sb;
  uint64_t* acc_full = b_empty + p.sb;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
  float* stat_s = (float*)(tmem_slot + 2);  // [2][nc]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, col0 = blockIdx.y * p.nc, n = blockIdx.z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.sa; i++) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.sb; i++) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&ta);
    tma_prefetch_desc(&tb);
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  if (p.has_stats)
    for (int i = threadIdx.x; i < 2 * p.nc; i += blockDim.x) stat_s[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- geometry of this CTA's rows
  const long long row0 = (long long)tile * p.m_cta;  // first flat row (CONV: within sample n)
  int mr_first = 0, a_row_base = 0;
  if (p.mode == IG_CONV) {
    mr_first = floordiv((int)row0 - p.Wp - 1, p.Wp);
    a_row_base = (int)row0 - mr_first * p.Wp;
  }

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer
    int sa = 0, pa = 0, sb = 0, pb = 0;
    const uint32_t a_tx = (p.mode == IG_CONV) ? (uint32_t)(p.nh_box * p.Wp * p.row_bytes)
                                              : (uint32_t)(p.mt * 128 * p.row_bytes);
    const uint32_t b_tx = (uint32_t)(p.nc * p.row_bytes);
    for (int kb = 0; kb < p.kblocks; kb++) {
      for (int g = 0; g < p.groups; g++) {
        mbar_wait(&a_empty[sa], pa ^ 1);
        mbar_expect_tx(&a_full[sa], a_tx);
        uint8_t* dst = a_s + (size_t)sa * p.slab_bytes;
        if (p.mode == IG_CONV) {
          tma_load_4d(dst, &ta, &a_full[sa], kb * p.kc, -1, mr_first + (g - 1) * p.H1, n);
        } else {
          for (int i = 0; i < p.mt; i++)
            tma_load_2d(dst + (size_t)i * 128 * p.row_bytes, &ta, &a_full[sa], kb * p.kc,
                        (int)(row0 + i * 128));
        }
        if (++sa == p.sa) { sa = 0; pa ^= 1; }
        for (int j = 0; j < p.tpg; j++) {
          const int tap = g * p.tpg + j;
          mbar_wait(&b_empty[sb], pb ^ 1);
          mbar_expect_tx(&b_full[sb], b_tx);
          tma_load_2d(b_s + (size_t)sb * p.b_bytes, &tb, &b_full[sb], kb * p.kc,
                      tap * p.b_rows_per_tap + col0);
          if (++sb == p.sb) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // =============================== MMA issuer
    int sa = 0, pa = 0, sb = 0, pb = 0;
    const uint32_t idesc = make_idesc(1, 128, (uint32_t)p.nc, 0, 0);
    const uint32_t layout = (p.row_bytes == 128) ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint64_t desc_hi = make_smem_desc(0, 16, 8 * p.row_bytes, layout);
    const int ksteps = p.row_bytes / 32;
    uint32_t accumulate = 0;
    for (int kb = 0; kb < p.kblocks; kb++) {
      for (int g = 0; g < p.groups; g++) {
        mbar_wait(&a_full[sa], pa);
        const uint32_t a_base = smem_u32(a_s + (size_t)sa * p.slab_bytes);
        for (int j = 0; j < p.tpg; j++) {
          int row_off = 0;
          if (p.mode == IG_CONV) row_off = a_row_base + (j / 3 - 1) * p.Wp + (j % 3 - 1);
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint32_t b_base = smem_u32(b_s + (size_t)sb * p.b_bytes);
          for (int mt = 0; mt < p.mt; mt++) {
            const uint32_t a_addr = a_base + (uint32_t)(row_off + mt * 128) * p.row_bytes;
#pragma unroll 4
            for (int ks = 0; ks < ksteps; ks++) {
              const uint64_t ad = desc_hi | (uint64_t)(((a_addr + ks * 32) >> 4) & 0x3FFF);
              const uint64_t bd = desc_hi | (uint64_t)(((b_base + ks * 32) >> 4) & 0x3FFF);
              umma_bf16(tmem + mt * p.nc, ad, bd, idesc, (ks > 0) ? 1u : accumulate);
            }
          }
          accumulate = 1;
          umma_commit(&b_empty[sb]);
          if (++sb == p.sb) { sb = 0; pb ^= 1; }
        }
        umma_commit(&a_empty[sa]);
        if (++sa == p.sa) { sa = 0; pa ^= 1; }
      }
    }
    umma_commit(acc_full);
  }
  __syncwarp();

  // =============================== epilogue (all four warps; warp w owns TMEM lanes 32w..32w+31)
  mbar_wait(acc_full, 0);
  tc_fence_after();
  const int t_idx = (p.out_mode == OUT_CONVT) ? col0 / p.cout_total : 0;
  const int co0 = (p.out_mode == OUT_CONVT) ? col0 % p.cout_total : col0;
  for (int mt = 0; mt < p.mt; mt++) {
    const long long r = row0 + mt * 128 + warp * 32 + lane;
    bool valid;
    long long off;  // element offset of this row's first output column
    if (p.out_mode == OUT_FLAT) {
      const int f = (int)r;
      const int mr = f / p.Wp, wq = f - mr * p.Wp;
      valid = (wq >= 1) && (mr < p.MR) && ((mr % p.H1) >= 1);
      off = (((long long)n * p.MR + mr) * p.W + (wq - 1)) * p.ldc + co0;
    } else if (p.out_mode == OUT_ROWS) {
      valid = r < p.rows_total;
      off = r * p.ldc + co0;
    } else {
      // coarse H-padded row r = ((n*D + d)*(H+1) + h')*W + w  ->  fine voxel (2d+i, 2h+j, 2w+k)
      valid = r < p.rows_total;
      long long q = r;
      const int w = (int)(q % p.ct_W); q /= p.ct_W;
      const int hp = (int)(q % (p.ct_H + 1)); q /= (p.ct_H + 1);
      const int d = (int)(q % p.ct_D);
      const long long nn = q / p.ct_D;
      valid = valid && hp >= 1;
      const int i = t_idx >> 2, j = (t_idx >> 1) & 1, k = t_idx & 1;
      const long long fd = 2 * d + i, fh = 2 * (hp - 1) + j + 1, fw = 2 * w + k;
      off = (((nn * (2 * p.ct_D) + fd) * (2 * p.ct_H + 1) + fh) * (2 * p.ct_W) + fw) * p.ldc + co0;
    }
    for (int c = 0; c < p.nc; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + mt * p.nc + c, v);
      tmem_ld_wait();
      float y[32];
#pragma unroll
      for (int i = 0; i < 32; i++) y[i] = __uint_as_float(v[i]);
      if (p.has_bias) {
#pragma unroll
        for (int i = 0; i < 32; i++) y[i] += __ldg(&p.bias[co0 + c + i]);
      }
      if (p.out_fp32) {
        if (valid) {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off + c);
#pragma unroll
          for (int i = 0; i < 8; i++) o[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
        }
      } else {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
          pk[i] = *reinterpret_cast<uint32_t*>(&h);
          // statistics are taken over exactly the values that are stored
          y[2 * i] = __low2float(h);
          y[2 * i + 1] = __high2float(h);
        }
        if (valid) {
          uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off + c);
#pragma unroll
          for (int i = 0; i < 4; i++) o[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
      }
      if (p.has_stats) {
        float s1[32], s2[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float t = valid ? y[i] : 0.f;
          s1[i] = t;
          s2[i] = t * t;
        }
        // transpose-reduce: after the loop lane L holds the column-(c+L) total in s[0]
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
          const bool up = (lane & s) != 0;
#pragma unroll
          for (int i = 0; i < s; i++) {
            const float send1 = up ? s1[i] : s1[i + s];
            const float keep1 = up ? s1[i + s] : s1[i];
            s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, s);
            const float send2 = up ? s2[i] : s2[i + s];
            const float keep2 = up ? s2[i + s] : s2[i];
            s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, s);
          }
        }
        atomicAdd(&stat_s[c + lane], s1[0]);
        atomicAdd(&stat_s[p.nc + c + lane], s2[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.has_stats) {
    double* st = p.stats + (p.stats_per_sample ? (size_t)n * p.cout_total * 2 : 0);
    for (int i = threadIdx.x; i < p.nc; i += blockDim.x) {
      atomicAdd(&st[(size_t)(co0 + i) * 2 + 0], (double)stat_s[i]);
      atomicAdd(&st[(size_t)(co0 + i) * 2 + 1], (double)stat_s[p.nc + i]);
    }
  }
  if (warp == 2) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------ host side
static int pow2_cols(int c) {
  int t = 32;
  while (t < c) t <<= 1;
  return t;
}

struct IgemmLaunch {
  IgemmParams p;
  CUtensorMap ta, tb;
  dim3 grid;
  size_t smem;
};

static int finish_and_launch(IgemmLaunch& L, cudaStream_t stream) {
  IgemmParams& p = L.p;
  p.tmem_cols = pow2_cols(p.mt * p.nc);
  if (p.tmem_cols > 512) return fail(PCRL_ERR_ARG, "igemm: mt*nc=%d exceeds TMEM", p.mt * p.nc);
  p.b_bytes = ((p.nc * p.row_bytes + 1023) / 1024) * 1024;
  // stage counts: fill what shared memory allows (<= 200 KB), at least 2 + 2
  const size_t budget = 200 * 1024;
  p.sa = 2;
  p.sb = 3;
  while ((size_t)p.sa * p.slab_bytes + (size_t)p.sb * p.b_bytes > budget && p.sb > 2) p.sb--;
  if ((size_t)p.sa * p.slab_bytes + (size_t)p.sb * p.b_bytes > budget)
    return fail(PCRL_ERR_ARG, "igemm: tile does not fit shared memory (slab %d B, b %d B)",
                p.slab_bytes, p.b_bytes);
  if (p.groups * p.kblocks >= 3 && (size_t)3 * p.slab_bytes + (size_t)p.sb * p.b_bytes <= budget) p.sa = 3;
  while ((size_t)p.sa * p.slab_bytes + (size_t)(p.sb + 1) * p.b_bytes <= budget && p.sb < 6) p.sb++;
  L.smem = (size_t)p.sa * p.slab_bytes + (size_t)p.sb * p.b_bytes + (2 * p.sa + 2 * p.sb + 1) * 8 +
           16 + 2 * p.nc * 4 + 1024;
  static size_t configured = 0;
  if (L.smem > configured) {
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(igemm_kmajor_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  igemm_kmajor_kernel<<<L.grid, 128, L.smem, stream>>>(L.ta, L.tb, p);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

// B operand map: packed weights [taps*rows_per_tap][K] bf16, box (kc, nc)
static int make_b_map(CUtensorMap* tb, const void* w, int K, long long rows, int kc, int nc) {
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)K * 2};
  uint32_t box[2] = {(uint32_t)kc, (uint32_t)nc};
  return encode_map(tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, str, box,
                    kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

// 3x3x3 convolution (forward or data gradient) on H-padded NDHWC bf16 activations.
//   x: [N][D][H+1][W][Cin], w: [27][Cout][Cin] bf16, y: [N][D][H+1][W][Cout] (bf16, or fp32)
//   stats (optional): [Cout][2] (or [N][Cout][2]) fp64, ACCUMULATED (caller zeroes)
int conv3d_k3_igemm(const void* x, const void* w, void* y, double* stats, int stats_per_sample,
                    int out_fp32, int N, int D, int H, int W, int Cin, int Cout,
                    cudaStream_t stream) {
  PCRL_REQUIRE(Cin % 32 == 0 && Cin >= 32, "conv3d_k3: Cin=%d must be a multiple of 32", Cin);
  PCRL_REQUIRE(Cin == 32 || Cin % 64 == 0, "conv3d_k3: Cin=%d must be 32 or a multiple of 64", Cin);
  PCRL_REQUIRE(Cout % 32 == 0, "conv3d_k3: Cout=%d must be a multiple of 32", Cout);
  PCRL_REQUIRE(W + 1 <= 256 && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_k3: bad dims");
  IgemmLaunch L;
  memset(&L.p, 0, sizeof(L.p));
  IgemmParams& p = L.p;
  p.mode = IG_CONV;
  p.W = W; p.Wp = W + 1; p.H1 = H + 1; p.D = D; p.MR = D * (H + 1);
  p.kc = (Cin == 32) ? 32 : 64;
  p.row_bytes = p.kc * 2;
  p.kblocks = Cin / p.kc;
  p.groups = 3; p.tpg = 9;
  p.nc = (Cout % 128 == 0) ? 128 : (Cout % 64 == 0 ? 64 : 32);
  const long long flat = (long long)p.MR * p.Wp;
  // rows per CTA: 256 when the sample is large enough, TMEM allows mt*nc <= 512
  p.mt = (flat >= 256) ? 2 : 1;
  if (p.nc <= 64 && flat >= 2048) p.mt = 4;
  p.m_cta = p.mt * 128;
  p.nh_box = (p.m_cta + 3 * p.Wp + 1 + p.Wp - 1) / p.Wp;
  if (p.nh_box > 256) return fail(PCRL_ERR_UNSUPPORTED, "conv3d_k3: W=%d too small for box", W);
  p.slab_bytes = ((p.nh_box * p.Wp * p.row_bytes + 1023) / 1024) * 1024;
  p.b_rows_per_tap = Cout;
  p.out_mode = OUT_FLAT; p.out_fp32 = out_fp32; p.ldc = Cout; p.cout_total = Cout;
  p.has_stats = stats != nullptr; p.stats_per_sample = stats_per_sample;
  p.out = y; p.stats = stats;
  uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)p.MR, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)p.MR * W * Cin * 2};
  uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)p.Wp, (uint32_t)p.nh_box, 1};
  int rc = encode_map(&L.ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box,
                      p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  rc = make_b_map(&L.tb, w, Cin, 27LL * Cout, p.kc, p.nc);
  if (rc) return rc;
  L.grid = dim3((unsigned)((flat + p.m_cta - 1) / p.m_cta), (unsigned)(Cout / p.nc), (unsigned)N);
  return finish_and_launch(L, stream);
}

// Plain GEMM  C[rows][cols] = A[rows][K] * B[cols][K]^T (+ bias[col]); A, B bf16 row-major.
// out_mode OUT_ROWS stores C row-major with leading dimension ldc; OUT_CONVT scatters the
// (tap, cout) columns of a ConvTranspose3d(k=2,s=2) to the fine H-padded NDHWC tensor.
int gemm_nt_igemm(const void* a, const void* b, void* c, const float* bias, long long rows, int K,
                  int cols, int ldc, int out_fp32, int out_mode, int ct_D, int ct_H, int ct_W,
                  int ct_cout, cudaStream_t stream) {
  PCRL_REQUIRE(K % 64 == 0, "gemm_nt: K=%d must be a multiple of 64", K);
  PCRL_REQUIRE(cols % 32 == 0, "gemm_nt: cols=%d must be a multiple of 32", cols);
  IgemmLaunch L;
  memset(&L.p, 0, sizeof(L.p));
  IgemmParams& p = L.p;
  p.mode = IG_PLAIN;
  p.kc = 64; p.row_bytes = 128; p.kblocks = K / 64; p.groups = 1; p.tpg = 1;
  p.nc = (cols % 128 == 0) ? 128 : (cols % 64 == 0 ? 64 : 32);
  if (out_mode == OUT_CONVT) {
    PCRL_REQUIRE(ct_cout % p.nc == 0, "convT: Cout=%d must be a multiple of %d", ct_cout, p.nc);
  }
  p.mt = rows >= 256 ? 2 : 1;
  p.m_cta = p.mt * 128;
  p.slab_bytes = p.m_cta * p.row_bytes;
  p.b_rows_per_tap = 0;
  p.rows_total = rows;
  p.out_mode = out_mode; p.out_fp32 = out_fp32; p.ldc = ldc;
  p.cout_total = (out_mode == OUT_CONVT) ? ct_cout : cols;
  p.ct_D = ct_D; p.ct_H = ct_H; p.ct_W = ct_W;
  p.has_bias = bias != nullptr; p.bias = bias; p.out = c;
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)K * 2};
  uint32_t box[2] = {64, 128};
  int rc = encode_map(&L.ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, dims, str, box,
                      CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_b_map(&L.tb, b, K, cols, 64, p.nc);
  if (rc) return rc;
  L.grid = dim3((unsigned)((rows + p.m_cta - 1) / p.m_cta), (unsigned)(cols / p.nc), 1);
  return finish_and_launch(L, stream);
}

}  // namespace pcrl
