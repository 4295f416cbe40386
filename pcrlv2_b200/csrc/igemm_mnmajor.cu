// MN-major implicit-GEMM on tcgen05 (sm_100a): weight gradients.
//
//   dW[tap][co][ci] += sum_{n, f} dY[n][f][co] * X[n][f + shift(tap)][ci]       (3x3x3 conv)
//   dW[p][q]        += sum_r A[r][p] * B[r][q]                                 (plain A^T B)
//
// Both operands keep the reduction index (the voxel / row) as their slab row and the channel as
// the contiguous dimension, i.e. they are "MN-major" for the tensor core: no transposes are
// materialised.  The reduction is split over CTAs (grid.x); every CTA adds its fp32 partial into
// dW with 16-byte vector reductions (red.global.add.v4.f32), so dW must be zeroed (or hold a
// running sum) on entry.  In CONV mode a CTA owns one (dz, dy) pair and computes the three dx
// taps from one X slab (row-shifted views), K-stages are `nrows` merged rows of one sample in the
// flat index space of common.cuh; tails of the stage buffers are zero so the K loop can run in
// whole 16-row steps.
// No halo rows are loaded for the dx taps: slab row 0 of every merged row is the virtual zero column
// (TMA out-of-bounds fill), so the dY view starts at slab row 1 and pairs dY row k+1 with X rows
// k+0 / k+1 / k+2 (dx = -1 / 0 / +1); the only X row outside the box that meets a non-zero dY row is
// the virtual column of the NEXT merged row, i.e. zero -- which is what the never-written tail holds.
// PAIR mode (Cout chunk of 64): an M = 64 MMA occupies the tensor pipe as long as an M = 128 one, so
// a CTA owns TWO (dz, dy) groups instead of one: the second half of the M = 128 operand is the same
// dY chunk loaded `delta` merged rows earlier (dY[f - delta] * X[f + off_a] = tap off_a + delta), the X
// slab is shared.  The nine groups become five CTA groups (the last one alone; its upper half is
// not stored), i.e. 1.8x fewer tensor-pipe cycles on the 64-channel layers.
// NPAIR mode (fp32 operands, 64-channel X chunk = two 32-channel smem chunks, so the dx taps cannot be
// stacked along N and a lone N = 64 MMA costs as much as N = 128): the N = 128 operand is the X chunk
// of TWO (dz,dy) groups, loaded as two slabs.  With PAIR on top ("quad") one MMA chain covers four
// groups {xg0, xg0+delta, xg1, xg1+delta}; the nine groups take 3 CTA groups instead of 5 (4+3+2 used).
// Grid order: blockIdx.x = (dz,dy) group fastest, then the (Cout, Cin) chunk pair; blockIdx.y = the
// reduction split.  CTAs that run at the same time therefore read the SAME K-stages (L2 hits instead
// of nine DRAM passes over X and dY).
// Reference: autograd of nn.Conv3d / nn.ConvTranspose3d / nn.Linear weights reached from
// train_3d.py:148 (loss.backward()).
#include "common.cuh"
#include "sm100.cuh"
#include <stdlib.h>

namespace pcrl {

template <int TF32>
__device__ __forceinline__ void umma_any(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (TF32) umma_tf32(d, ad, bd, idesc, acc);   // compile-time kind: see igemm_kmajor.cu
  else umma_bf16(d, ad, bd, idesc, acc);
}

enum { WG_CONV = 0, WG_PLAIN = 1 };

struct WgradParams {
  int mode;
  int W, Wp, H1, MR, nsamples;
  int nrows;          // merged rows per K-stage (CONV) / rows per K-stage (PLAIN)
  int kr;             // reduction rows per stage: nrows*Wp (CONV) / nrows (PLAIN)
  int ksteps;         // ceil(kr/16)
  int stages_per_sample, total_stages, stages_per_cta;
  int mc, nc;         // M (dY / A channels) and N (X / B channels) per CTA
  int a_row_bytes, b_row_bytes;          // 128 (64-channel chunks) or 64 (32-channel chunk)
  int a_chunks, b_chunks;                // 64-channel chunks in mc / nc
  int a_chunk_bytes, b_chunk_bytes;      // bytes of one chunk region in smem (1024-aligned)
  int b_box_rows;                        // rows written by the X TMA box
  int ntaps;                             // taps per CTA (3 in CONV, 1 in PLAIN)
  int stages, tmem_cols;
  int rot_stages;                        // CONV: stage rotation per dz plane (see the producer)
  int nacc;                              // accumulator sets (1 or 2) that k-steps alternate between
  int tf32, krows;                       // fp32 operands / kind::tf32; reduction rows per MMA (16 or 8)
  int chunk_ch;                          // channels per 128-byte chunk row (64 bf16 / 32 fp32)
  int stack_dx;                          // CONV: one MMA of N = 3*nc covers the three dx taps
  int m_chunks_total;                    // Cout / mch
  int pair;                              // CONV: two (dz,dy) groups per CTA stacked along M (see header)
  int ngroups;                           // CTA groups over (dz,dy): 9, 5 in PAIR or NPAIR mode, 3 in both, 1 in PLAIN mode
  int npair;                             // CONV: two X groups stacked along N (see header)
  int b_views;                           // X slabs per stage: 2 in NPAIR mode, else 1
  int ncm;                               // MMA N without dx stacking: nc * b_views
  int mch;                               // dW rows (output channels) per m-chunk: mc, or 64 in PAIR mode
  int cout, cin;                         // leading dims of dW: [tap][cout][cin]
  long long rows_total;
  float* dw;
};

template <int TF32>
__global__ void __launch_bounds__(128)
igemm_mnmajor_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                     const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int a_stage_bytes = p.a_chunks * p.a_chunk_bytes;
  const int b_stage_bytes = p.b_chunks * p.b_views * p.b_chunk_bytes;
  const int stage_bytes = a_stage_bytes + b_stage_bytes;
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* acc_full = empty + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunk = blockIdx.y;
  // (dz,dy) group(s) of this CTA in CONV mode: ga for M rows 0..63 (all rows without PAIR), gb for 64..127
  const int gidx = (int)blockIdx.x % p.ngroups, mn = (int)blockIdx.x / p.ngroups;
  const int mchunk = mn % p.m_chunks_total, nchunk = mn / p.m_chunks_total;
  // xg[u]: group the X slab u is positioned for; delta: merged rows the second dY view lags;
  // tapg[v][u]: (dz,dy) group (0..8) accumulator block (dY view v, X slab u) holds, -1 = unused
  int xg[2], delta = 0, tapg[2][2] = {{-1, -1}, {-1, -1}};
  if (p.pair && p.npair) {
    if (gidx == 0) { xg[0] = 0; xg[1] = 3; delta = 1; tapg[0][0] = 0; tapg[1][0] = 1; tapg[0][1] = 3; tapg[1][1] = 4; }
    else if (gidx == 1) { xg[0] = 6; xg[1] = 2; delta = 1; tapg[0][0] = 6; tapg[1][0] = 7; tapg[0][1] = 2; }
    else { xg[0] = 5; xg[1] = 8; tapg[0][0] = 5; tapg[0][1] = 8; }
  } else if (p.pair) {
    const int ga = 2 * gidx, gb = min(ga + 1, 8);
    xg[0] = xg[1] = ga;
    delta = (gb / 3 - ga / 3) * p.H1 + (gb % 3 - ga % 3);
    tapg[0][0] = ga;
    if (gb != ga) tapg[1][0] = gb;
  } else if (p.npair) {
    xg[0] = 2 * gidx; xg[1] = min(xg[0] + 1, 8);
    tapg[0][0] = xg[0];
    if (xg[1] != xg[0]) tapg[0][1] = xg[1];
  } else {
    xg[0] = xg[1] = gidx;
    tapg[0][0] = gidx;
  }
  const int dzo = xg[0] / 3 - 1;

  // zero all stage buffers once: tails beyond what TMA writes must be finite (dY tails: zero)
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = p.stages * stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&ta);
    tma_prefetch_desc(&tb);
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  fence_proxy_async();  // generic-proxy zero fill must be visible to the async proxy (TMA / MMA)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int s_begin = kchunk * p.stages_per_cta;
  const int s_end = min(s_begin + p.stages_per_cta, p.total_stages);

  if (warp == 0) {
    // producer: the whole warp loops (operands stay warp-uniform), one elected lane issues
    int st = 0, ph = 0;
    const uint32_t tx = (uint32_t)(p.a_chunks * p.kr * p.a_row_bytes +
                                   p.b_chunks * p.b_views * p.b_box_rows * p.b_row_bytes);
    for (int s = s_begin; s < s_end; s++) {
      mbar_wait(&empty[st], ph ^ 1);
      uint8_t* a_dst = smem + (size_t)st * stage_bytes;
      uint8_t* b_dst = a_dst + a_stage_bytes;
      if (elect_one()) {
        mbar_expect_tx(&full[st], tx);
        if (p.mode == WG_CONV) {
          // stage rotation (rot_stages = dz * H1 / nrows when X is the larger operand): the groups of
          // the three dz planes then read the SAME X rows at the same time (they differ in dY rows
          // instead), so X is fetched from DRAM once; any permutation of the stages is a valid order
          int se = s - dzo * p.rot_stages;
          if (se < 0) se += p.total_stages;
          else if (se >= p.total_stages) se -= p.total_stages;
          const int n = se / p.stages_per_sample;
          const int mr0 = (se % p.stages_per_sample) * p.nrows;
          const int cpv = p.a_chunks >> p.pair;            // channel chunks per dY view
          for (int c = 0; c < p.a_chunks; c++) {
            const int view = c / cpv, cc = c - view * cpv;
            tma_load_4d(a_dst + (size_t)c * p.a_chunk_bytes, &ta, &full[st],
                        mchunk * p.mch + cc * p.chunk_ch, -1, mr0 - view * delta, n);
          }
          for (int u = 0; u < p.b_views; u++) {
            const int xoff = (xg[u] / 3 - 1) * p.H1 + (xg[u] % 3 - 1);
            for (int c = 0; c < p.b_chunks; c++)
              tma_load_4d(b_dst + (size_t)(u * p.b_chunks + c) * p.b_chunk_bytes, &tb, &full[st],
                          nchunk * p.nc + c * p.chunk_ch, -1, mr0 + xoff, n);
          }
        } else {
          const int r0 = s * p.nrows;
          for (int c = 0; c < p.a_chunks; c++)
            tma_load_2d(a_dst + (size_t)c * p.a_chunk_bytes, &ta, &full[st], mchunk * p.mc + c * p.chunk_ch, r0);
          for (int c = 0; c < p.b_chunks; c++)
            tma_load_2d(b_dst + (size_t)c * p.b_chunk_bytes, &tb, &full[st], nchunk * p.nc + c * p.chunk_ch, r0);
        }
      }
      __syncwarp();
      if (++st == p.stages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    int st = 0, ph = 0;
    const uint32_t fmt = TF32 ? 2u : 1u;
    const uint32_t idesc = make_idesc(fmt, (uint32_t)p.mc, (uint32_t)p.ncm, 1, 1);
    // MN-major layouts: 16-bit operands use the 128B/64B swizzle (K atom = 8 rows); fp32 (tf32)
    // operands need the 32-byte-atom variant (K atom = 4 rows) -- profiles/r01_umma_probe.md
    const uint32_t a_lay = TF32 ? LAYOUT_SW128_B32 : (p.a_row_bytes == 128 ? LAYOUT_SW128 : LAYOUT_SW64);
    const uint32_t b_lay = TF32 ? LAYOUT_SW128_B32 : (p.b_row_bytes == 128 ? LAYOUT_SW128 : LAYOUT_SW64);
    const uint32_t a_sbo = TF32 ? 512u : 8u * p.a_row_bytes, b_sbo = TF32 ? 512u : 8u * p.b_row_bytes;
    const uint64_t a_hi = make_smem_desc(0, p.a_chunk_bytes, a_sbo, a_lay);
    const uint64_t b_hi = make_smem_desc(0, p.b_chunk_bytes, b_sbo, b_lay);
    const uint32_t a_step = (uint32_t)(p.krows * p.a_row_bytes) >> 4, b_step = (uint32_t)(p.krows * p.b_row_bytes) >> 4;
    const uint32_t idesc_stack = make_idesc(fmt, (uint32_t)p.mc, (uint32_t)(3 * p.nc), 1, 1);
    const uint64_t b_stack_hi = make_smem_desc(0, p.b_row_bytes, b_sbo, b_lay);
    // Back-to-back MMAs into the SAME accumulator run at ~3/4 of the rate of MMAs that rotate over
    // several (tools/umma_probe: N = 128, 1554 vs 2026 TFLOP/s), so consecutive MMAs never share one:
    // the three dx taps are interleaved inside the k-step loop, and where TMEM has room (nacc = 2)
    // even / odd k-steps use two accumulator sets that the epilogue adds.
    const uint32_t set_cols = (uint32_t)(p.ntaps * p.ncm);
    const uint32_t b_row16 = (uint32_t)p.b_row_bytes >> 4;
    uint32_t kq = 0;                                  // k-steps issued so far by this CTA
    for (int s = s_begin; s < s_end; s++) {
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t a_slab = smem_u32(smem + (size_t)st * stage_bytes);
      const uint32_t b_base = a_slab + a_stage_bytes;
      // CONV: the dY view starts one row into the slab (see header)
      const uint32_t a_base = a_slab + (p.mode == WG_CONV ? (uint32_t)p.a_row_bytes : 0u);
      if (elect_one()) {
        uint64_t ad = a_hi | (uint64_t)((a_base >> 4) & 0x3FFF);
        if (p.stack_dx) {
          // the three dx taps are the same X slab shifted by one row each: with a single channel
          // chunk per tap they form ONE MN-major operand of N = 3*nc whose chunk stride (LBO) is one
          // slab row, so a single MMA fills the three accumulators (N = 96 / 192 instead of 3 x 32 / 64)
          uint64_t bd = b_stack_hi | (uint64_t)((b_base >> 4) & 0x3FFF);
          for (int ks = 0; ks < p.ksteps; ks++, kq++) {
            const uint32_t set = kq & (uint32_t)(p.nacc - 1);
            umma_any<TF32>(tmem + set * set_cols, ad, bd, idesc_stack, kq >= (uint32_t)p.nacc ? 1u : 0u);
            ad += a_step;
            bd += b_step;
          }
        } else if (p.ntaps == 3) {
          uint64_t bd = b_hi | (uint64_t)((b_base >> 4) & 0x3FFF);
          for (int ks = 0; ks < p.ksteps; ks++, kq++) {
            const uint32_t d = tmem + (kq & (uint32_t)(p.nacc - 1)) * set_cols;
            const uint32_t acc = kq >= (uint32_t)p.nacc ? 1u : 0u;
            umma_any<TF32>(d, ad, bd, idesc, acc);
            umma_any<TF32>(d + (uint32_t)p.ncm, ad, bd + b_row16, idesc, acc);
            umma_any<TF32>(d + 2u * (uint32_t)p.ncm, ad, bd + 2u * b_row16, idesc, acc);
            ad += a_step;
            bd += b_step;
          }
        } else {
          uint64_t bd = b_hi | (uint64_t)((b_base >> 4) & 0x3FFF);
          for (int ks = 0; ks < p.ksteps; ks++, kq++) {
            const uint32_t set = kq & (uint32_t)(p.nacc - 1);
            umma_any<TF32>(tmem + set * set_cols, ad, bd, idesc, kq >= (uint32_t)p.nacc ? 1u : 0u);
            ad += a_step;
            bd += b_step;
          }
        }
        umma_commit(&empty[st]);
      } else {
        kq += (uint32_t)p.ksteps;
      }
      __syncwarp();
      if (++st == p.stages) { st = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  }
  __syncwarp();

  if (s_begin < s_end) {
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // M = 128: accumulator row i sits in TMEM lane i.  M = 64: lane (i/16)*32 + i%16.
    int row;
    bool row_ok;
    if (p.mc == 128) { row = warp * 32 + lane; row_ok = true; }
    else { row = warp * 16 + lane; row_ok = lane < 16; }
    // the second accumulator set exists only if this CTA issued at least two k-steps
    const bool two_sets = p.nacc == 2 && (long long)(s_end - s_begin) * p.ksteps >= 2;
    const int vw = (p.pair && row >= 64) ? 1 : 0;        // dY view of this row (warp-uniform)
    const int orow = p.pair ? (row & 63) : row;
    for (int t = 0; t < p.ntaps; t++) {
      for (int u = 0; u < p.b_views; u++) {
        const int g = (p.mode == WG_CONV) ? tapg[vw][u] : 0;
        if (g < 0) continue;                               // unused block of a PAIR / NPAIR chain
        const int tap = (p.mode == WG_CONV) ? g * 3 + t : 0;
        float* dst = p.dw + ((size_t)tap * p.cout + (size_t)mchunk * p.mch + orow) * p.cin +
                     (size_t)nchunk * p.nc;
        for (int c = 0; c < p.nc; c += 32) {
          const uint32_t col = (uint32_t)(t * p.ncm + u * p.nc + c);
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + col, v);
          tmem_ld_wait();
          if (two_sets) {
            uint32_t v2[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(p.ntaps * p.ncm) + col, v2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 8; i++)
              red_add_v4(dst + c + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

static int launch_wgrad(WgradParams& p, const CUtensorMap& ta, const CUtensorMap& tb, dim3 grid,
                        cudaStream_t stream) {
  if (p.b_views < 1) { p.b_views = 1; p.ncm = p.nc; }
  const int stage_bytes = p.a_chunks * p.a_chunk_bytes + p.b_chunks * p.b_views * p.b_chunk_bytes;
  if (p.stages < 2) {
    p.stages = 3;
    while (p.stages > 2 && (size_t)p.stages * stage_bytes > 200 * 1024) p.stages--;
  }
  if ((size_t)p.stages * stage_bytes > 225 * 1024)
    return fail(PCRL_ERR_ARG, "wgrad: stage of %d bytes does not fit shared memory", stage_bytes);
  p.nacc = (2 * p.ntaps * p.ncm <= 512 && !getenv("PCRL_WGRAD_NACC1")) ? 2 : 1;
  int cols = p.nacc * p.ntaps * p.ncm, t = 32;
  while (t < cols) t <<= 1;
  p.tmem_cols = t;
  const size_t smem = (size_t)p.stages * stage_bytes + (2 * p.stages + 1) * 8 + 16 + 1024;
  if (getenv("PCRL_WGRAD_VERBOSE"))
    fprintf(stderr, "[wgrad] mode %d tf32 %d mc %d nc %d nrows %d kr %d ksteps %d stages %d stage_bytes %d grid %u x %u pair %d npair %d stack %d\n",
            p.mode, p.tf32, p.mc, p.nc, p.nrows, p.kr, p.ksteps, p.stages, stage_bytes, grid.x, grid.y, p.pair, p.npair, p.stack_dx);
  static bool configured = false;
  if (!configured) {
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(igemm_mnmajor_kernel<0>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(igemm_mnmajor_kernel<1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (p.tf32) igemm_mnmajor_kernel<1><<<grid, 128, smem, stream>>>(ta, tb, p);
  else igemm_mnmajor_kernel<0><<<grid, 128, smem, stream>>>(ta, tb, p);
  PCRL_CHECK_LAUNCH();
  return PCRL_OK;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// dW[27][Cout][Cin] (fp32) += conv3d_k3 weight gradient.
//   dy: [N][D][H+1][W][Cout] bf16 (pad rows zero), x: [N][D][H+1][W][Cin] bf16 (pad rows zero)
int conv3d_k3_wgrad_igemm(const void* dy, const void* x, float* dw, int N, int D, int H, int W,
                          int Cin, int Cout, cudaStream_t stream, int dtype) {
  const int tf32 = dtype != PCRL_DTYPE_BF16;
  PCRL_REQUIRE(Cout % 64 == 0, "conv3d_k3_wgrad: Cout=%d must be a multiple of 64", Cout);
  PCRL_REQUIRE(Cin == 32 || Cin % 64 == 0, "conv3d_k3_wgrad: Cin=%d must be 32 or a multiple of 64", Cin);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.mode = WG_CONV;
  p.tf32 = tf32;
  const int elt = tf32 ? 4 : 2;
  p.krows = 32 / elt;
  p.chunk_ch = 128 / elt;
  p.W = W; p.Wp = W + 1; p.H1 = H + 1; p.MR = D * (H + 1); p.nsamples = N;
  p.pair = (Cout % 128 != 0 && !getenv("PCRL_WGRAD_NOPAIR")) ? 1 : 0;
  p.mc = (Cout % 128 == 0 || p.pair) ? 128 : 64;
  p.mch = p.pair ? 64 : p.mc;
  p.nc = (Cin % 128 == 0) ? 128 : (Cin % 64 == 0 ? 64 : 32);
  p.a_row_bytes = 128; p.a_chunks = p.mc / p.chunk_ch;
  if (!tf32 && p.nc == 32) { p.b_row_bytes = 64; p.b_chunks = 1; }
  else { p.b_row_bytes = 128; p.b_chunks = p.nc / p.chunk_ch; }
  p.npair = (tf32 && p.nc == 64 && p.b_chunks == 2 && !getenv("PCRL_WGRAD_NONPAIR")) ? 1 : 0;
  p.b_views = p.npair ? 2 : 1;
  p.ncm = p.nc * p.b_views;
  // Merged rows per K-stage and pipeline depth.  Measured on B200 (tools/sweep_wgrad.py,
  // profiles/r02p_sweep_wgrad.txt): what matters first is a LONG stage -- a stage of >= ~160 reduction
  // rows amortises the per-stage barrier round trip and the zero-padded last k-step (kr is rarely a
  // multiple of the MMA K) -- and only then the depth: two stages of the largest tiling that fits
  // beat three stages of a shorter one by 10-30 %, three stages win by a few % once they are long too.
  const int target = 272;                                   // reduction rows per stage worth having
  int nrows0 = target / p.Wp;
  if (nrows0 < 1) nrows0 = 1;
  if (nrows0 > p.MR) nrows0 = p.MR;
  if (nrows0 > 256) nrows0 = 256;
  if (const char* e = getenv("PCRL_WGRAD_NROWS")) { if (atoi(e) > 0 && atoi(e) <= p.MR && atoi(e) <= 256) nrows0 = atoi(e); }
  auto largest_fit = [&](int want) {
    for (int nr = nrows0; nr >= 1; nr--) {
      const int ks = (nr * p.Wp + p.krows - 1) / p.krows;
      const int ab = round_up((ks * p.krows + 1) * p.a_row_bytes, 1024);
      const int bb = round_up((ks * p.krows + 3) * p.b_row_bytes, 1024);
      if ((size_t)want * (p.a_chunks * ab + p.b_chunks * p.b_views * bb) <= 225 * 1024) return nr;
    }
    return 0;
  };
  int chosen = 0;
  if (const char* e = getenv("PCRL_WGRAD_STAGES")) {
    if (atoi(e) >= 2 && atoi(e) <= 4) { p.stages = atoi(e); chosen = largest_fit(p.stages); }
  }
  if (!chosen) {
    const int nr3 = largest_fit(3), nr2 = largest_fit(2);
    if (nr3 > 0 && (nr3 == nr2 || nr3 * p.Wp >= 160)) { chosen = nr3; p.stages = 3; }
    else { chosen = nr2; p.stages = 2; }
  }
  if (!chosen) return fail(PCRL_ERR_ARG, "conv3d_k3_wgrad: no stage tiling fits shared memory");
  p.nrows = chosen;
  p.kr = p.nrows * p.Wp;
  p.ksteps = (p.kr + p.krows - 1) / p.krows;
  // PAIR: the second dY view lags by up to H1 - 2 merged rows (groups (dz=-1,dy=+1) | (dz=0,dy=-1)),
  // so the reduction range is extended by that much (rows outside the sample are TMA zero fill)
  const int k_extra = p.pair ? (p.npair ? 1 : (p.H1 - 2 > 1 ? p.H1 - 2 : 1)) : 0;
  p.stages_per_sample = (p.MR + k_extra + p.nrows - 1) / p.nrows;
  p.total_stages = p.stages_per_sample * N;
  // (measured on B200: +3 % with fp32 operands, where the unrotated working set overflows L2; -2 % with bf16)
  p.rot_stages = (tf32 && Cin >= Cout && !p.npair && !getenv("PCRL_WGRAD_NOROT")) ? p.H1 / p.nrows : 0;
  if (p.rot_stages >= p.total_stages) p.rot_stages = 0;
  p.a_chunk_bytes = round_up((p.ksteps * p.krows + 1) * p.a_row_bytes, 1024);
  p.b_box_rows = p.nrows * p.Wp;
  p.b_chunk_bytes = round_up((p.ksteps * p.krows + 3) * p.b_row_bytes, 1024);
  p.ntaps = 3;
  p.stack_dx = (p.b_chunks == 1 && !getenv("PCRL_WGRAD_NOSTACK")) ? 1 : 0;
  p.m_chunks_total = Cout / p.mch;
  p.cout = Cout; p.cin = Cin; p.dw = dw;
  // split the reduction so that the grid has a few waves
  p.ngroups = (p.pair && p.npair) ? 3 : ((p.pair || p.npair) ? 5 : 9);
  const int other = p.ngroups * (Cout / p.mch) * (Cin / p.nc);
  int kchunks = (4 * num_sms() + other - 1) / other;
  if (kchunks > p.total_stages) kchunks = p.total_stages;
  if (kchunks < 1) kchunks = 1;
  p.stages_per_cta = (p.total_stages + kchunks - 1) / kchunks;
  kchunks = (p.total_stages + p.stages_per_cta - 1) / p.stages_per_cta;
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)p.MR, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cout * elt, (uint64_t)W * Cout * elt, (uint64_t)p.MR * W * Cout * elt};
    uint32_t box[4] = {(uint32_t)p.chunk_ch, (uint32_t)p.Wp, (uint32_t)p.nrows, 1};
    int rc = encode_map(&ta, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy, dims, str, box,
                        tf32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)p.MR, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * elt, (uint64_t)W * Cin * elt, (uint64_t)p.MR * W * Cin * elt};
    uint32_t box[4] = {(uint32_t)(p.b_row_bytes / elt), (uint32_t)p.Wp, (uint32_t)p.nrows, 1};
    int rc = encode_map(&tb, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box,
                        tf32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                             : (p.b_row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B));
    if (rc) return rc;
  }
  dim3 grid((unsigned)other, (unsigned)kchunks, 1);
  return launch_wgrad(p, ta, tb, grid, stream);
}

// dW[P][Q] (fp32, leading dimension Q) += A[rows][P]^T * B[rows][Q]; A, B bf16 row-major.
int gemm_tn_igemm(const void* a, const void* b, float* dw, long long rows, int P, int Q,
                  cudaStream_t stream, int dtype) {
  const int tf32 = dtype != PCRL_DTYPE_BF16;
  PCRL_REQUIRE(P % 64 == 0 && (Q % 64 == 0 || Q == 32), "gemm_tn: P=%d (multiple of 64), Q=%d (32 or multiple of 64)", P, Q);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.mode = WG_PLAIN;
  p.tf32 = tf32;
  const int elt = tf32 ? 4 : 2;
  p.krows = 32 / elt;
  p.chunk_ch = 128 / elt;
  p.nrows = tf32 ? 64 : 128; p.kr = p.nrows; p.ksteps = p.nrows / p.krows;
  p.total_stages = (int)((rows + p.nrows - 1) / p.nrows);
  p.stages_per_sample = p.total_stages;
  p.mc = (P % 128 == 0) ? 128 : 64;
  p.nc = (Q % 128 == 0) ? 128 : (Q % 64 == 0 ? 64 : 32);
  p.a_row_bytes = 128; p.a_chunks = p.mc / p.chunk_ch;
  if (!tf32 && p.nc == 32) { p.b_row_bytes = 64; p.b_chunks = 1; }
  else { p.b_row_bytes = 128; p.b_chunks = p.nc / p.chunk_ch; }
  p.a_chunk_bytes = p.nrows * 128; p.b_chunk_bytes = p.nrows * p.b_row_bytes;
  p.b_box_rows = p.nrows;
  p.ntaps = 1;
  p.m_chunks_total = P / p.mc;
  p.mch = p.mc;
  p.ngroups = 1;
  p.cout = P; p.cin = Q; p.dw = dw; p.rows_total = rows;
  const int other = (P / p.mc) * (Q / p.nc);
  int kchunks = (2 * num_sms() + other - 1) / other;
  if (kchunks > p.total_stages) kchunks = p.total_stages;
  if (kchunks < 1) kchunks = 1;
  p.stages_per_cta = (p.total_stages + kchunks - 1) / kchunks;
  kchunks = (p.total_stages + p.stages_per_cta - 1) / p.stages_per_cta;
  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {(uint64_t)P, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)P * elt};
    uint32_t box[2] = {(uint32_t)p.chunk_ch, (uint32_t)p.nrows};
    int rc = encode_map(&ta, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, dims, str, box,
                        tf32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Q, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)Q * elt};
    uint32_t box[2] = {(uint32_t)(p.b_row_bytes / elt), (uint32_t)p.nrows};
    int rc = encode_map(&tb, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, dims, str, box,
                        tf32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                             : (p.b_row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B));
    if (rc) return rc;
  }
  dim3 grid((unsigned)other, (unsigned)kchunks, 1);
  return launch_wgrad(p, ta, tb, grid, stream);
}

}  // namespace pcrl
