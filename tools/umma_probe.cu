// tcgen05 / TMA probe for sm_100a: checks the descriptor conventions the conv kernels rely on
// (row-shifted operand views inside swizzled slabs, MN-major operands, tf32, 5-D TMA halo
// fill) against a CPU reference, and measures SS-mode MMA issue throughput per tile shape.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu
// Run:   umma_probe <test> [args...]   (one test per process; a trap cannot poison the others)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include "../pcrlv2_b200/csrc/sm100.cuh"

using namespace pcrl;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA_ERROR %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(3);                                                                     \
    }                                                                              \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(3); }
  return (EncodeTiledFn)fn;
}
static CUtensorMap make_map(CUtensorMapDataType dt, int rank, void* ptr, const uint64_t* dims,
                            const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box,
                            CUtensorMapSwizzle sw) {
  static EncodeTiledFn enc = get_encode();
  CUtensorMap m;
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; i++) s[i] = strides_bytes[i];
  CUresult r = enc(&m, dt, rank, ptr, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(3); }
  return m;
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
static float tf32_trunc(float x) {  // tensor cores read the top 19 bits of an fp32 operand
  uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x;
}
static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

// ------------------------------------------------------------------------------------------
struct KParams {
  int n, kblocks, ra, shift, bo_mode, layout, row_bytes, is_tf32, mtiles;
};

// K-major probe: A slab [ra rows][row_bytes] per k-block (TMA, swizzled or interleaved),
// B [n rows][row_bytes]; D[r][j] = sum_k A[r + shift][k] * B[j][k] for r < 128*mtiles.
__global__ void __launch_bounds__(128) probe_kmajor(const __grid_constant__ CUtensorMap ta,
                                                   const __grid_constant__ CUtensorMap tb,
                                                   KParams p, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int a_bytes = ((p.ra * p.row_bytes + 1023) / 1024) * 1024;
  const int b_bytes = ((p.n * p.row_bytes + 1023) / 1024) * 1024;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + (size_t)p.kblocks * a_bytes;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int elts_per_row = p.row_bytes / (p.is_tf32 ? 4 : 2);
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_full, (uint32_t)(p.kblocks * (p.ra + p.n) * p.row_bytes));
    for (int kb = 0; kb < p.kblocks; kb++) {
      if (p.layout == LAYOUT_NONE) {
        tma_load_3d(a_s + (size_t)kb * a_bytes, &ta, &bar_full, 0, 0, kb * (p.row_bytes / 16));
        tma_load_3d(b_s + (size_t)kb * b_bytes, &tb, &bar_full, 0, 0, kb * (p.row_bytes / 16));
      } else {
        tma_load_2d(a_s + (size_t)kb * a_bytes, &ta, &bar_full, kb * elts_per_row, 0);
        tma_load_2d(b_s + (size_t)kb * b_bytes, &tb, &bar_full, kb * elts_per_row, 0);
      }
    }
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(p.is_tf32 ? 2 : 1, 128, p.n, 0, 0);
    const int ksteps = p.row_bytes / 32;
    for (int mt = 0; mt < p.mtiles; mt++) {
      uint32_t acc = 0;
      for (int kb = 0; kb < p.kblocks; kb++) {
        for (int ks = 0; ks < ksteps; ks++) {
          uint32_t a_addr, b_addr;
          uint64_t ad, bd;
          if (p.layout == LAYOUT_NONE) {
            // [chunk16B][row][16B]: K-adjacent core matrices ra*16 (n*16) apart, M-adjacent 128 B
            a_addr = smem_u32(a_s + (size_t)kb * a_bytes) + (p.shift + mt * 128) * 16 +
                     ks * 2 * p.ra * 16;
            b_addr = smem_u32(b_s + (size_t)kb * b_bytes) + ks * 2 * p.n * 16;
            ad = make_smem_desc(a_addr, p.ra * 16, 128, LAYOUT_NONE);
            bd = make_smem_desc(b_addr, p.n * 16, 128, LAYOUT_NONE);
          } else {
            a_addr = smem_u32(a_s + (size_t)kb * a_bytes) + (p.shift + mt * 128) * p.row_bytes +
                     ks * 32;
            b_addr = smem_u32(b_s + (size_t)kb * b_bytes) + ks * 32;
            uint32_t bo = p.bo_mode ? ((a_addr >> 7) & 7) : 0;
            ad = make_smem_desc(a_addr, 16, 8 * p.row_bytes, p.layout, bo);
            bd = make_smem_desc(b_addr, 16, 8 * p.row_bytes, p.layout, 0);
          }
          if (p.is_tf32) umma_tf32(tmem + mt * p.n, ad, bd, idesc, acc);
          else umma_bf16(tmem + mt * p.n, ad, bd, idesc, acc);
          acc = 1;
        }
      }
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int mt = 0; mt < p.mtiles; mt++) {
    for (int c = 0; c < p.n; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + mt * p.n + c, v);
      tmem_ld_wait();
      float* o = out + (size_t)(mt * 128 + warp * 32 + (threadIdx.x & 31)) * p.n + c;
      for (int i = 0; i < 32; i++) o[i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int run_kmajor(KParams p) {
  const int esz = p.is_tf32 ? 4 : 2;
  const int epr = p.row_bytes / esz;  // elements per row per k-block
  const int K = epr * p.kblocks;
  const int M = 128 * p.mtiles;
  if (p.ra < M + p.shift) { printf("bad ra\n"); return 2; }
  std::vector<float> A((size_t)p.ra * K), B((size_t)p.n * K);
  for (auto& x : A) x = p.is_tf32 ? tf32_trunc(frand()) : bf16_round(frand());
  for (auto& x : B) x = p.is_tf32 ? tf32_trunc(frand()) : bf16_round(frand());
  void *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, A.size() * esz)); CK(cudaMalloc(&dB, B.size() * esz));
  CK(cudaMalloc(&dO, (size_t)M * p.n * 4));
  if (p.is_tf32) {
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  } else {
    std::vector<__nv_bfloat16> a16(A.size()), b16(B.size());
    for (size_t i = 0; i < A.size(); i++) a16[i] = __float2bfloat16(A[i]);
    for (size_t i = 0; i < B.size(); i++) b16[i] = __float2bfloat16(B[i]);
    CK(cudaMemcpy(dA, a16.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, b16.data(), B.size() * 2, cudaMemcpyHostToDevice));
  }
  CUtensorMapDataType dt = p.is_tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ta, tb;
  if (p.layout == LAYOUT_NONE) {
    uint64_t da[3] = {(uint64_t)(16 / esz), (uint64_t)p.ra, (uint64_t)(K * esz / 16)};
    uint64_t sa[2] = {(uint64_t)K * esz, 16};
    uint32_t ba[3] = {(uint32_t)(16 / esz), (uint32_t)p.ra, (uint32_t)(p.row_bytes / 16)};
    ta = make_map(dt, 3, dA, da, sa, ba, CU_TENSOR_MAP_SWIZZLE_NONE);
    uint64_t db[3] = {(uint64_t)(16 / esz), (uint64_t)p.n, (uint64_t)(K * esz / 16)};
    uint32_t bb[3] = {(uint32_t)(16 / esz), (uint32_t)p.n, (uint32_t)(p.row_bytes / 16)};
    tb = make_map(dt, 3, dB, db, sa, bb, CU_TENSOR_MAP_SWIZZLE_NONE);
  } else {
    CUtensorMapSwizzle sw = p.layout == LAYOUT_SW128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : p.layout == LAYOUT_SW64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                    : CU_TENSOR_MAP_SWIZZLE_32B;
    uint64_t da[2] = {(uint64_t)K, (uint64_t)p.ra};
    uint64_t sa[1] = {(uint64_t)K * esz};
    uint32_t ba[2] = {(uint32_t)epr, (uint32_t)p.ra};
    ta = make_map(dt, 2, dA, da, sa, ba, sw);
    uint64_t db[2] = {(uint64_t)K, (uint64_t)p.n};
    uint32_t bb[2] = {(uint32_t)epr, (uint32_t)p.n};
    tb = make_map(dt, 2, dB, db, sa, bb, sw);
  }
  size_t smem = (size_t)p.kblocks * ((((size_t)p.ra * p.row_bytes + 1023) / 1024) * 1024 +
                                     (((size_t)p.n * p.row_bytes + 1023) / 1024) * 1024) + 2048;
  CK(cudaFuncSetAttribute(probe_kmajor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kmajor<<<1, 128, smem>>>(ta, tb, p, dO);
  CK(cudaDeviceSynchronize());
  std::vector<float> O((size_t)M * p.n);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < M; r++)
    for (int j = 0; j < p.n; j++) {
      double ref = 0;
      for (int k = 0; k < K; k++) ref += (double)A[(size_t)(r + p.shift) * K + k] * B[(size_t)j * K + k];
      maxerr = fmax(maxerr, fabs(ref - O[(size_t)r * p.n + j]));
      maxref = fmax(maxref, fabs(ref));
    }
  bool ok = maxerr <= 1e-3 * fmax(maxref, 1.0);
  printf("RESULT kmajor layout=%d rowB=%d tf32=%d n=%d kb=%d ra=%d mt=%d shift=%d bo=%d : maxerr=%.3e maxref=%.3e %s\n",
         p.layout, p.row_bytes, p.is_tf32, p.n, p.kblocks, p.ra, p.mtiles, p.shift, p.bo_mode,
         maxerr, maxref, ok ? "PASS" : "FAIL");
  return ok ? 0 : 1;
}

// ------------------------------------------------------------------------------------------
struct MNParams { int m, n, kr, sa, sb, ksteps, tf32, sbo, layout, tmasw; };

// MN-major probe (weight-gradient shape): A[kr rows][m] and B[kr rows][n] are row slabs whose
// rows are the reduction index; D[i][j] = sum_{k < 16*ksteps} A[k + sa][i] * B[k + sb][j].
__global__ void __launch_bounds__(128) probe_mnmajor(const __grid_constant__ CUtensorMap ta,
                                                    const __grid_constant__ CUtensorMap tb,
                                                    MNParams p, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int chunk_bytes = ((p.kr * 128 + 1023) / 1024) * 1024;  // one 128-byte column chunk
  const int cch = p.tf32 ? 32 : 64;                              // channels per chunk
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + (size_t)(p.m / cch) * chunk_bytes;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_full, (uint32_t)((p.m / cch + p.n / cch) * p.kr * 128));
    for (int c = 0; c < p.m / cch; c++) tma_load_2d(a_s + (size_t)c * chunk_bytes, &ta, &bar_full, c * cch, 0);
    for (int c = 0; c < p.n / cch; c++) tma_load_2d(b_s + (size_t)c * chunk_bytes, &tb, &bar_full, c * cch, 0);
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(p.tf32 ? 2 : 1, p.m, p.n, 1, 1);
    const int kr = p.tf32 ? 8 : 16;
    for (int ks = 0; ks < p.ksteps; ks++) {
      uint32_t a_addr = smem_u32(a_s) + (p.sa + ks * kr) * 128;
      uint32_t b_addr = smem_u32(b_s) + (p.sb + ks * kr) * 128;
      uint64_t ad = make_smem_desc(a_addr, chunk_bytes, p.sbo, p.layout);
      uint64_t bd = make_smem_desc(b_addr, chunk_bytes, p.sbo, p.layout);
      if (p.tf32) umma_tf32(tmem, ad, bd, idesc, ks > 0);
      else umma_bf16(tmem, ad, bd, idesc, ks > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  if (p.m == 128 || warp < 2) {
    // M=64 accumulators occupy lanes 0..15 of each 32-lane quadrant? (checked on the host side:
    // we dump all 128 lanes and let the host search the layout)
  }
  for (int c = 0; c < p.n; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    float* o = out + (size_t)(warp * 32 + (threadIdx.x & 31)) * p.n + c;
    for (int i = 0; i < 32; i++) o[i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int run_mnmajor(MNParams p) {
  std::vector<float> A((size_t)p.kr * p.m), B((size_t)p.kr * p.n);
  for (auto& x : A) x = p.tf32 ? tf32_trunc(frand()) : bf16_round(frand());
  for (auto& x : B) x = p.tf32 ? tf32_trunc(frand()) : bf16_round(frand());
  const int esz = p.tf32 ? 4 : 2, cch = p.tf32 ? 32 : 64;
  void *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, A.size() * esz)); CK(cudaMalloc(&dB, B.size() * esz));
  CK(cudaMalloc(&dO, (size_t)128 * p.n * 4));
  CK(cudaMemset(dO, 0, (size_t)128 * p.n * 4));
  if (p.tf32) {
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  } else {
    std::vector<__nv_bfloat16> a16(A.size()), b16(B.size());
    for (size_t i = 0; i < A.size(); i++) a16[i] = __float2bfloat16(A[i]);
    for (size_t i = 0; i < B.size(); i++) b16[i] = __float2bfloat16(B[i]);
    CK(cudaMemcpy(dA, a16.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, b16.data(), B.size() * 2, cudaMemcpyHostToDevice));
  }
  CUtensorMapDataType dt = p.tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  uint64_t da[2] = {(uint64_t)p.m, (uint64_t)p.kr}; uint64_t sa[1] = {(uint64_t)p.m * esz};
  uint32_t ba[2] = {(uint32_t)cch, (uint32_t)p.kr};
  CUtensorMapSwizzle tsw = (CUtensorMapSwizzle)p.tmasw;
  CUtensorMap ta = make_map(dt, 2, dA, da, sa, ba, tsw);
  uint64_t db[2] = {(uint64_t)p.n, (uint64_t)p.kr}; uint64_t sb[1] = {(uint64_t)p.n * esz};
  CUtensorMap tb = make_map(dt, 2, dB, db, sb, ba, tsw);
  size_t chunk = (((size_t)p.kr * 128 + 1023) / 1024) * 1024;
  size_t smem = (p.m / cch + p.n / cch) * chunk + 2048;
  CK(cudaFuncSetAttribute(probe_mnmajor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_mnmajor<<<1, 128, smem>>>(ta, tb, p, dO);
  CK(cudaDeviceSynchronize());
  std::vector<float> O((size_t)128 * p.n);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  const int K = (p.tf32 ? 8 : 16) * p.ksteps;
  double maxerr = 0, maxref = 0;
  // M=128: row i -> lane i.  M=64: try lane = i (lanes 0..63) and the 16-per-quadrant layout.
  double maxerr_alt = 0;
  for (int i = 0; i < p.m; i++)
    for (int j = 0; j < p.n; j++) {
      double ref = 0;
      for (int k = 0; k < K; k++) ref += (double)A[(size_t)(k + p.sa) * p.m + i] * B[(size_t)(k + p.sb) * p.n + j];
      maxerr = fmax(maxerr, fabs(ref - O[(size_t)i * p.n + j]));
      int lane_alt = (i / 16) * 32 + (i % 16);
      maxerr_alt = fmax(maxerr_alt, fabs(ref - O[(size_t)lane_alt * p.n + j]));
      maxref = fmax(maxref, fabs(ref));
    }
  bool ok = maxerr <= 1e-3 * fmax(maxref, 1.0);
  bool ok_alt = maxerr_alt <= 1e-3 * fmax(maxref, 1.0);
  printf("RESULT mnmajor layout=%d tmasw=%d tf32=%d sbo=%d m=%d n=%d kr=%d ksteps=%d sa=%d sb=%d : maxerr=%.3e alt=%.3e maxref=%.3e %s%s\n",
         p.layout, p.tmasw, p.tf32, p.sbo, p.m, p.n, p.kr, p.ksteps, p.sa, p.sb, maxerr, maxerr_alt, maxref, ok ? "PASS" : "FAIL",
         ok_alt ? " (alt-lane-layout PASS)" : "");
  return ok ? 0 : 1;
}

// ------------------------------------------------------------------------------------------
// 5-D TMA halo probe: X[N][D][H][W][C=64] bf16, box (64, W+1, bh+2, 1, 1) at (0,-1,h0-1,d,n).
__global__ void probe_halo(const __grid_constant__ CUtensorMap tx, int rows, int w0, int h0, int d0,
                           int n0, __nv_bfloat16* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    mbar_expect_tx(&bar, rows * 128);
    tma_load_5d(smem, &tx, &bar, 0, w0, h0, d0, n0);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  // de-swizzle on the way out: 16-byte chunk c of row r sits at chunk (c ^ (r & 7))
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    int r = i / 8, c = i % 8;
    const uint4 v = *reinterpret_cast<const uint4*>(smem + r * 128 + ((c ^ (r & 7)) * 16));
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + (size_t)r * 128 + c * 16) = v;
  }
}

static int run_halo() {
  const int N = 2, D = 3, H = 6, W = 4, C = 64, bh = 4;
  std::vector<float> X((size_t)N * D * H * W * C);
  for (auto& x : X) x = bf16_round(frand());
  std::vector<__nv_bfloat16> x16(X.size());
  for (size_t i = 0; i < X.size(); i++) x16[i] = __float2bfloat16(X[i]);
  __nv_bfloat16 *dX, *dO;
  CK(cudaMalloc(&dX, X.size() * 2));
  CK(cudaMemcpy(dX, x16.data(), X.size() * 2, cudaMemcpyHostToDevice));
  const int rows = (W + 1) * (bh + 2);
  CK(cudaMalloc(&dO, rows * 128));
  uint64_t dims[5] = {C, W, H, D, N};
  uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
  uint32_t box[5] = {64, W + 1, bh + 2, 1, 1};
  CUtensorMap tx = make_map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dX, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  int fails = 0;
  const int cases[4][3] = {{0, 1, 0}, {3, 2, 1}, {0, -1, 0}, {3, 3, 1}};  // h0, d, n
  for (auto& cs : cases) {
    int h0 = cs[0], d = cs[1], n = cs[2];
    CK(cudaFuncSetAttribute(probe_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 128 + 2048));
    probe_halo<<<1, 128, rows * 128 + 2048>>>(tx, rows, -1, h0 - 1, d, n, dO);
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> O((size_t)rows * 64);
    CK(cudaMemcpy(O.data(), dO, O.size() * 2, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int hh = 0; hh < bh + 2; hh++)
      for (int ww = 0; ww < W + 1; ww++)
        for (int c = 0; c < C; c++) {
          int h = h0 - 1 + hh, w = ww - 1;
          float ref = 0;
          if (h >= 0 && h < H && w >= 0 && w < W && d >= 0 && d < D)
            ref = X[((((size_t)n * D + d) * H + h) * W + w) * C + c];
          maxerr = fmax(maxerr, fabs(ref - __bfloat162float(O[(size_t)(hh * (W + 1) + ww) * 64 + c])));
        }
    printf("RESULT halo5d h0=%d d=%d n=%d : maxerr=%.3e %s\n", h0, d, n, maxerr, maxerr == 0 ? "PASS" : "FAIL");
    fails += maxerr != 0;
  }
  return fails;
}

// ------------------------------------------------------------------------------------------
// Issue-throughput probe: every CTA runs `iters` rounds of (taps x mtiles x 4) MMAs over a
// resident slab, operand A shifted by `tap_rows * tap` rows per tap.  No loads in the loop.
struct PerfParams { int n, mtiles, taps, tap_rows, iters, layout; };
__global__ void __launch_bounds__(128) perf_mma(PerfParams p, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const int a_rows = 128 * p.mtiles + p.taps * p.tap_rows + 8;
  for (int i = threadIdx.x; i < (a_rows + p.n) * 128 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(1, 128, p.n, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + ((a_rows * 128 + 1023) / 1024) * 1024;
    for (int it = 0; it < p.iters; it++)
      for (int t = 0; t < p.taps; t++)
        for (int mt = 0; mt < p.mtiles; mt++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            uint64_t ad, bd;
            if (p.layout == LAYOUT_NONE) {
              ad = make_smem_desc(a0 + (mt * 128 + t * p.tap_rows) * 16 + ks * 2 * a_rows * 16, a_rows * 16, 128, LAYOUT_NONE);
              bd = make_smem_desc(b0 + ks * 2 * p.n * 16, p.n * 16, 128, LAYOUT_NONE);
            } else {
              ad = make_smem_desc(a0 + (mt * 128 + t * p.tap_rows) * 128 + ks * 32, 16, 1024, LAYOUT_SW128);
              bd = make_smem_desc(b0 + ks * 32, 16, 1024, LAYOUT_SW128);
            }
            umma_bf16(tmem + mt * p.n, ad, bd, idesc, 1);
          }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  if (v[0] == 0x12345678u) sink[0] = 1.f;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int run_perf(PerfParams p) {
  float* sink; CK(cudaMalloc(&sink, 4));
  int a_rows = 128 * p.mtiles + p.taps * p.tap_rows + 8;
  size_t smem = ((a_rows * 128 + 1023) / 1024) * 1024 + p.n * 128 + 4096;
  CK(cudaFuncSetAttribute(perf_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    perf_mma<<<sms, 128, smem>>>(p, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 128 * p.n * 16 * 4.0 * p.mtiles * p.taps * (double)p.iters * sms;
    if (rep == 2)
      printf("RESULT perf layout=%d n=%d mtiles=%d taps=%d tap_rows=%d iters=%d sms=%d : %.3f ms  %.1f TFLOP/s\n",
             p.layout, p.n, p.mtiles, p.taps, p.tap_rows, p.iters, sms, ms, flops / ms * 1e-9);
  }
  return 0;
}


// ------------------------------------------------------------------------------------------
// Mixed probe: the MMA loop of perf_mma (N columns, mtiles x taps x 4 MMAs per round, `iters`
// rounds) running concurrently with a bulk-copy ingest stream (`nloads` chunks of `chunk` bytes
// per CTA through a `stages`-deep ring, read from a `gbytes`-sized global buffer).  Either side
// can be switched off (iters = 0 / nloads = 0) to get the two isolated rates.
struct MixParams { int n, mtiles, taps, iters, nloads, chunk, stages; long long gbytes; int same;
                   int commit_every, tap_rows, pollers, fence_every; };
__global__ void __launch_bounds__(192) mix_kernel(MixParams p, const uint8_t* g, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma, full[16], empty[16], dummy[4], never;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const int a_rows = 128 * p.mtiles + p.taps * p.tap_rows + 8;
  const uint32_t op_bytes = ((a_rows * 128 + 1023) / 1024) * 1024 + p.n * 128;
  uint8_t* ring = smem + ((op_bytes + 1023) / 1024) * 1024;
  for (int i = threadIdx.x; i < (int)op_bytes / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    for (int i = 0; i < p.stages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 4; i++) mbar_init(&dummy[i], 1);
    mbar_init(&never, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc(1, 128, p.n, 0, 0);
      const uint32_t a0 = smem_u32(smem), b0 = a0 + ((a_rows * 128 + 1023) / 1024) * 1024;
      int cnt = 0, dm = 0;
      for (int it = 0; it < p.iters; it++)
        for (int t = 0; t < p.taps; t++) {
          if (p.fence_every) tc_fence_after();
          for (int mt = 0; mt < p.mtiles; mt++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
              uint64_t ad = make_smem_desc(a0 + (mt * 128 + t * p.tap_rows) * 128 + ks * 32, 16, 1024, LAYOUT_SW128);
              uint64_t bd = make_smem_desc(b0 + ks * 32, 16, 1024, LAYOUT_SW128);
              umma_bf16(tmem + mt * p.n, ad, bd, idesc, 1);
            }
            if (p.commit_every && ++cnt == p.commit_every) { cnt = 0; umma_commit(&dummy[dm]); dm = (dm + 1) & 3; }
          }
        }
      umma_commit(&bar_mma);
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    if (threadIdx.x == 0) mbar_arrive(&never);
  } else if (warp >= 1 && warp <= 3 && warp <= p.pollers) {
    // pollers: whole warps spinning on a barrier that flips only when the MMAs are done
    mbar_wait(&never, 0);
  } else if (warp == 4) {
    // loader
    if ((threadIdx.x & 31) == 0) {
      int s = 0, ph = 0;
      const long long nch = p.gbytes / p.chunk;
      long long c = p.same ? 0 : ((long long)blockIdx.x * 7919) % nch;
      for (int i = 0; i < p.nloads; i++) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], (uint32_t)p.chunk);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(ring + (size_t)s * p.chunk)),
            "l"(g + c * p.chunk), "r"((uint32_t)p.chunk), "r"(smem_u32(&full[s]))
            : "memory");
        if (++c == nch) c = 0;
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    // drain: release each stage as soon as it has landed
    if ((threadIdx.x & 31) == 0) {
      int s = 0, ph = 0;
      for (int i = 0; i < p.nloads; i++) {
        mbar_wait(&full[s], ph);
        mbar_arrive(&empty[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t v[32];
    tmem_ld32(tmem, v);
    tmem_ld_wait();
    if (v[0] == 0x12345678u) sink[0] = 1.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int run_mix(MixParams p) {
  float* sink; CK(cudaMalloc(&sink, 4));
  uint8_t* g; CK(cudaMalloc(&g, p.gbytes)); CK(cudaMemset(g, 0, p.gbytes));
  int a_rows = 128 * p.mtiles + p.taps * p.tap_rows + 8;
  size_t smem = ((a_rows * 128 + 1023) / 1024) * 1024 + p.n * 128 + 2048 + (size_t)p.stages * p.chunk + 1024;
  CK(cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    mix_kernel<<<sms, 192, smem>>>(p, g, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 128 * p.n * 16 * 4.0 * p.mtiles * p.taps * (double)p.iters * sms;
    double bytes = (double)p.nloads * p.chunk * sms;
    if (rep == 2)
      printf("RESULT mix n=%d mt=%d iters=%d nloads=%d chunk=%d stages=%d gMB=%.1f same=%d ce=%d tr=%d poll=%d fence=%d : %.3f ms  %.1f TFLOP/s  %.2f TB/s  %.1f B/clk/SM@%dMHz\n",
             p.n, p.mtiles, p.iters, p.nloads, p.chunk, p.stages, p.gbytes / 1048576.0, p.same, p.commit_every, p.tap_rows, p.pollers, p.fence_every, ms, flops / ms * 1e-9,
             bytes / ms * 1e-9, bytes / sms / (ms * 1e-3 * khz * 1e3), khz / 1000);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// Lean issue probe: descriptors precomputed, inner loops fully unrolled -- measures the floor
// of clk per tcgen05.mma for a given N when the issuing thread does nothing else.
template <int MT>
__global__ void __launch_bounds__(128) lean_kernel(int n, int iters, int whole_warp, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const int a_rows = 128 * MT + 8;
  for (int i = threadIdx.x; i < (a_rows + n) * 128 / 4 + 256; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(1, 128, n, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + ((a_rows * 128 + 1023) / 1024) * 1024;
    const uint64_t ad = make_smem_desc(a0, 16, 1024, LAYOUT_SW128);
    const uint64_t bd = make_smem_desc(b0, 16, 1024, LAYOUT_SW128);
    if (whole_warp) {
      __shared__ uint64_t dummy[8];
      if (threadIdx.x == 0) { for (int i = 0; i < 8; i++) mbar_init(&dummy[i], 1); fence_barrier_init(); }
      __syncwarp();
      int sb = 0;
      for (int it = 0; it < iters; it++) {
        if (whole_warp & 4) mbar_wait(&dummy[(sb + 4) & 7], 1);   // parity of the phase before the first: completes at once
        if (whole_warp & 8) tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int ks = 0; ks < 4; ks++)
              umma_bf16(tmem + mt * n, ad + mt * 1024 + 2 * ks, bd + 2 * ks, idesc, 1);
          if (whole_warp & 2) umma_commit(&dummy[sb]);
        }
        __syncwarp();
        sb = (sb + 1) & 3;
      }
      if (elect_one()) umma_commit(&bar_mma);
    } else if (threadIdx.x == 0) {
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            umma_bf16(tmem + mt * n, ad + mt * 1024 + 2 * ks, bd + 2 * ks, idesc, 1);
      }
      umma_commit(&bar_mma);
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(tmem, v);
    tmem_ld_wait();
    if (v[0] == 0x12345678u) sink[0] = 1.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int run_lean(int n, int mt, int iters, int whole_warp) {
  float* sink; CK(cudaMalloc(&sink, 4));
  size_t smem = (size_t)(128 * mt + 8 + n) * 128 + 4096;
  auto k = mt == 1 ? lean_kernel<1> : (mt == 2 ? lean_kernel<2> : lean_kernel<4>);
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    k<<<sms, 128, smem>>>(n, iters, whole_warp, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double nmma = 4.0 * mt * iters;
    double flops = 2.0 * 128 * n * 16 * nmma * sms;
    if (rep == 2)
      printf("RESULT lean n=%d mt=%d iters=%d warp=%d : %.3f ms  %.1f TFLOP/s  %.1f clk/MMA@%dMHz\n", n, mt, iters,
             whole_warp, ms, flops / ms * 1e-9, ms * 1e-3 * khz * 1e3 / nmma, khz / 1000);
  }
  return 0;
}

// Dual-issue probe: `nw` warps each issue their own MMA stream (own accumulator columns), with the
// same per-8-MMA overhead options as lean (flags: 2 commit, 4 wait, 8 fence, 16 prefetched wait).
__global__ void __launch_bounds__(128) dual_kernel(int n, int iters, int flags, int nw, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma[4], dummy[4][8];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const int a_rows = 128 * 2 + 8;
  for (int i = threadIdx.x; i < (a_rows + n) * 128 / 4 + 256; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; w++) { mbar_init(&bar_mma[w], 1); for (int i = 0; i < 8; i++) mbar_init(&dummy[w][i], 1); }
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s + warp * n;
  if (warp < nw) {
    const uint32_t idesc = make_idesc(1, 128, n, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + ((a_rows * 128 + 1023) / 1024) * 1024;
    const uint64_t ad = make_smem_desc(a0, 16, 1024, LAYOUT_SW128) + (warp & 1) * 1024;
    const uint64_t bd = make_smem_desc(b0, 16, 1024, LAYOUT_SW128);
    int sb = 0;
    bool ready = true;
    for (int it = 0; it < iters; it++) {
      if (flags & 4) mbar_wait(&dummy[warp][(sb + 4) & 7], 1);
      if (flags & 16) { if (!ready) mbar_wait(&dummy[warp][(sb + 4) & 7], 1); ready = mbar_try_wait(&dummy[warp][(sb + 5) & 7], 1); }
      if (flags & 8) tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            umma_bf16(tmem, ad + 2 * ks, bd + 2 * ks, idesc, 1);
        if (flags & 2) umma_commit(&dummy[warp][sb]);
      }
      __syncwarp();
      sb = (sb + 1) & 3;
    }
    if (elect_one()) umma_commit(&bar_mma[warp]);
    __syncwarp();
    mbar_wait(&bar_mma[warp], 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(tmem_base_s, v);
    tmem_ld_wait();
    if (v[0] == 0x12345678u) sink[0] = 1.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

static int run_dual(int n, int iters, int flags, int nw) {
  float* sink; CK(cudaMalloc(&sink, 4));
  size_t smem = (size_t)(128 * 2 + 8 + n) * 128 + 4096;
  CK(cudaFuncSetAttribute(dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    dual_kernel<<<sms, 128, smem>>>(n, iters, flags, nw, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double nmma = 8.0 * iters * nw;
    double flops = 2.0 * 128 * n * 16 * nmma * sms;
    if (rep == 2)
      printf("RESULT dual n=%d iters=%d flags=%d warps=%d : %.3f ms  %.1f TFLOP/s  %.1f clk/MMA@%dMHz\n", n, iters,
             flags, nw, ms, flops / ms * 1e-9, ms * 1e-3 * khz * 1e3 / nmma, khz / 1000);
  }
  return 0;
}


// ------------------------------------------------------------------------------------------
// cta2: tcgen05.mma.cta_group::2 (CTA pair, M = 256).  Each CTA of the pair holds 128 rows of A
// (plus `shift` spare rows for the row-shifted view the conv kernels use) and n/2 rows of B; the
// leader issues the MMAs, each CTA's TMEM receives its own 128 accumulator rows x n columns.
// Checks the result against a CPU reference for both candidate B-half assignments and measures the
// issue rate with operands resident in shared memory (what a CTA pair buys: every SM reads only
// half of B per MMA, i.e. less shared-memory bandwidth per flop than two independent M = 128 MMAs).
struct C2Params { int n, kblocks, shift, iters, ra; };

__device__ __forceinline__ void umma_bf16_cta2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cta2(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe_cta2(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, C2Params p,
           float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int nh = p.n / 2;
  const int a_bytes = ((p.ra * 128 + 1023) / 1024) * 1024;
  const int b_bytes = ((nh * 128 + 1023) / 1024) * 1024;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + (size_t)p.kblocks * a_bytes;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_full, (uint32_t)(p.kblocks * (p.ra + nh) * 128));
    for (int kb = 0; kb < p.kblocks; kb++) {
      tma_load_2d(a_s + (size_t)kb * a_bytes, &ta, &bar_full, kb * 64, (int)rank * p.ra);
      tma_load_2d(b_s + (size_t)kb * b_bytes, &tb, &bar_full, kb * 64, (int)rank * nh);
    }
    mbar_wait(&bar_full, 0);
  }
  __syncthreads();
  cluster_sync_all();   // both CTAs' operands are in shared memory
  tc_fence_after();
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(1, 256, p.n, 0, 0);
    for (int it = 0; it < p.iters; it++) {
      uint32_t acc = 0;
      for (int kb = 0; kb < p.kblocks; kb++)
        for (int ks = 0; ks < 4; ks++) {
          const uint32_t a_addr = smem_u32(a_s + (size_t)kb * a_bytes) + p.shift * 128 + ks * 32;
          const uint32_t b_addr = smem_u32(b_s + (size_t)kb * b_bytes) + ks * 32;
          umma_bf16_cta2(tmem, make_smem_desc(a_addr, 16, 1024, LAYOUT_SW128, 0),
                         make_smem_desc(b_addr, 16, 1024, LAYOUT_SW128, 0), idesc, acc);
          acc = 1;
        }
    }
    umma_commit_cta2(&bar_mma, 3);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  if (pair == 0) {
    for (int c = 0; c < p.n; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      float* o = out + (size_t)(rank * 128 + warp * 32 + (threadIdx.x & 31)) * p.n + c;
      for (int i = 0; i < 32; i++) o[i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static int run_cta2(C2Params p) {
  const int K = 64 * p.kblocks, nh = p.n / 2;
  p.ra = 128 + ((p.shift + 7) / 8) * 8;
  std::vector<float> A((size_t)2 * p.ra * K), B((size_t)p.n * K);
  for (auto& x : A) x = bf16_round(frand());
  for (auto& x : B) x = bf16_round(frand());
  std::vector<__nv_bfloat16> a16(A.size()), b16(B.size());
  for (size_t i = 0; i < A.size(); i++) a16[i] = __float2bfloat16(A[i]);
  for (size_t i = 0; i < B.size(); i++) b16[i] = __float2bfloat16(B[i]);
  void *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, a16.size() * 2)); CK(cudaMalloc(&dB, b16.size() * 2));
  CK(cudaMalloc(&dO, (size_t)256 * p.n * 4));
  CK(cudaMemcpy(dA, a16.data(), a16.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, b16.data(), b16.size() * 2, cudaMemcpyHostToDevice));
  uint64_t da[2] = {(uint64_t)K, (uint64_t)(2 * p.ra)}, sa[1] = {(uint64_t)K * 2};
  uint32_t ba[2] = {64, (uint32_t)p.ra};
  CUtensorMap ta = make_map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, da, sa, ba, CU_TENSOR_MAP_SWIZZLE_128B);
  uint64_t db[2] = {(uint64_t)K, (uint64_t)p.n};
  uint32_t bb[2] = {64, (uint32_t)nh};
  CUtensorMap tb = make_map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, db, sa, bb, CU_TENSOR_MAP_SWIZZLE_128B);
  size_t smem = (size_t)p.kblocks * ((((size_t)p.ra * 128 + 1023) / 1024) * 1024 + (((size_t)nh * 128 + 1023) / 1024) * 1024) + 2048;
  CK(cudaFuncSetAttribute(probe_cta2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  C2Params q = p; q.iters = 1;
  probe_cta2<<<2, 128, smem>>>(ta, tb, q, dO);
  CK(cudaDeviceSynchronize());
  std::vector<float> O((size_t)256 * p.n);
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  // hypothesis 0: column j < n/2 comes from CTA 0's B rows, j >= n/2 from CTA 1's (B row j of the global matrix)
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 256; m++)
    for (int j = 0; j < p.n; j++) {
      const int arow = (m / 128) * p.ra + (m % 128) + p.shift;
      double ref = 0;
      for (int k = 0; k < K; k++) ref += (double)A[(size_t)arow * K + k] * B[(size_t)j * K + k];
      maxerr = fmax(maxerr, fabs(ref - O[(size_t)m * p.n + j]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("RESULT cta2 n=%d kb=%d shift=%d : maxerr=%.3e maxref=%.3e %s\n", p.n, p.kblocks, p.shift, maxerr, maxref,
         maxerr <= 1e-3 * fmax(maxref, 1.0) ? "OK" : "MISMATCH");
  // issue rate: every pair of SMs runs `iters` passes over the resident operands
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = (sms / 2) * 2;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    probe_cta2<<<grid, 128, smem>>>(ta, tb, p, dO);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double nmma = (double)p.iters * p.kblocks * 4;
    const double flops = 2.0 * 256 * p.n * 16 * nmma * (grid / 2);
    if (rep == 2)
      printf("RESULT cta2 rate n=%d iters=%d pairs=%d : %.3f ms  %.1f TFLOP/s  %.1f clk/MMA@%dMHz\n", p.n, p.iters,
             grid / 2, ms, flops / ms * 1e-9, ms * 1e-3 * khz * 1e3 / nmma, khz / 1000);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage\n"); return 2; }
  std::string t = argv[1];
  auto I = [&](int i, int d) { return argc > i ? atoi(argv[i]) : d; };
  if (t == "kmajor") {
    // kmajor layout row_bytes tf32 n kblocks ra mtiles shift bo_mode
    KParams p;
    p.layout = I(2, 2); p.row_bytes = I(3, 128); p.is_tf32 = I(4, 0); p.n = I(5, 64);
    p.kblocks = I(6, 2); p.ra = I(7, 256); p.mtiles = I(8, 1); p.shift = I(9, 0); p.bo_mode = I(10, 0);
    return run_kmajor(p);
  } else if (t == "mnmajor") {
    MNParams p; p.m = I(2, 128); p.n = I(3, 64); p.kr = I(4, 128); p.ksteps = I(5, 4); p.sa = I(6, 0); p.sb = I(7, 0); p.tf32 = I(8, 0); p.sbo = I(9, 1024); p.layout = I(10, 2); p.tmasw = I(11, 3);
    return run_mnmajor(p);
  } else if (t == "halo") {
    return run_halo();
  } else if (t == "cta2") {
    // cta2 n kblocks shift iters
    C2Params p; p.n = I(2, 256); p.kblocks = I(3, 2); p.shift = I(4, 0); p.iters = I(5, 2000);
    return run_cta2(p);
  } else if (t == "dual") {
    return run_dual(I(2, 128), I(3, 4000), I(4, 0), I(5, 2));
  } else if (t == "lean") {
    return run_lean(I(2, 128), I(3, 2), I(4, 4000), I(5, 0));
  } else if (t == "mix") {
    // mix n mtiles taps iters nloads chunk stages gMB same
    MixParams p; p.n = I(2, 128); p.mtiles = I(3, 2); p.taps = I(4, 9); p.iters = I(5, 100); p.nloads = I(6, 1000);
    p.chunk = I(7, 16384); p.stages = I(8, 8); p.gbytes = (long long)I(9, 64) * 1048576 / (I(9, 64) < 0 ? 1 : 1); p.same = I(10, 0);
    if (I(9, 64) == 0) p.gbytes = 524288;
    p.commit_every = I(11, 0); p.tap_rows = I(12, 0); p.pollers = I(13, 0); p.fence_every = I(14, 0);
    return run_mix(p);
  } else if (t == "perf") {
    PerfParams p; p.n = I(2, 64); p.mtiles = I(3, 1); p.taps = I(4, 9); p.tap_rows = I(5, 0); p.iters = I(6, 200); p.layout = I(7, 2);
    return run_perf(p);
  }
  return 2;
}
