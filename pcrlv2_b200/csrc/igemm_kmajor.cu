// K-major implicit-GEMM on tcgen05 (sm_100a):   OUT[row][col] = sum_tap sum_k A_tap[row][k] * B[tap][col][k]
//
// One persistent, warp-specialised kernel serves
//   * 3x3x3 convolution forward  (A = activations, B = packed filter)    -- LUConv, reference
//     models/pcrlv2_model_3d.py:9,33 (nn.Conv3d(k=3, padding=1))
//   * its data gradient          (A = dY, B = mirrored / transposed filter)
//   * ConvTranspose3d(k=2,s=2) forward as a GEMM with a scatter epilogue (reference :52,64)
//   * plain row-major GEMMs (ConvTranspose data gradient).
//
// CONV mode works on the per-sample flat index space of the H-padded NDHWC layout (common.cuh).
// A tile is m_cta consecutive flat rows of P "segments" spaced one plane (PL = (H+1)*(W+1) rows)
// apart, i.e. the same (h,w) window of P consecutive output planes.  For every (k-block, input
// plane q = -1..P) the producer TMA-loads ONE slab of whole merged rows (box (kc, W+1, nh, 1),
// halo zero-filled by out-of-bounds handling); the nine (dy,dx) taps are row-shifted views of
// that slab (descriptor start address + shift*row_bytes, verified on hardware,
// profiles/r01_umma_probe.md), and the up-to-three output planes an input plane feeds
// (dz = +1, 0, -1) are stacked along the MMA N dimension: one MMA of N = cnt*nc columns per
// (tap, 128-row block, 16-channel step).  Stacking lifts N from 64 to 128/192 for the 64-channel
// layers, where a single N=64 MMA only reaches ~46 % of the tensor-pipe rate.
// Weights stream through a second TMA ring ([cnt*nc x kc] per tap).
//
// Warp roles (256 threads): warp 0 = slab TMA producer, warp 6 = filter TMA producer, warps 1 and 7 =
// MMA issuers, warps 2..5 = epilogue.  TWO issuers because the issuing thread is the critical
// resource: a tcgen05.mma of N <= 128 occupies the pipe for only 51..69 clk and the pipe takes the
// next one only from a thread that is ready to issue, so every barrier wait / commit / branch of a
// single issuer is dead tensor time (measured: tools/umma_probe "lean" and "dual",
// profiles/r01_issue_probe.md).  Issuer w owns the 128-row blocks mt = w, w+2, .. of the tile (its
// own accumulator columns); while one issuer is in its per-tap bookkeeping the other one's MMAs run.
// Accumulators live in TMEM, double-buffered when 2*mt*P*nc <= 512 columns so that the epilogue
// of tile i (tcgen05.ld -> bias -> bf16/fp32 store, per-channel sum / sum-of-squares for the
// following BatchNorm / InstanceNorm) overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "sm100.cuh"
#include <stdlib.h>

namespace pcrl {

enum { IG_CONV = 0, IG_PLAIN = 1 };
enum { OUT_FLAT = 0, OUT_ROWS = 1, OUT_CONVT = 2, OUT_ROWS_T = 3, OUT_UNSHUFFLE = 4 };

struct IgemmParams {
  int mode;
  // geometry (CONV)
  int W, Wp, H1, D, MR, PL;   // H1 = H+1, MR = D*H1 merged rows per sample, PL = H1*Wp
  int nh_box;
  // tiling
  int m_cta, mt, P, nc, kc, row_bytes, kblocks, tpg;
  int slab_bytes, b_bytes, sa, sb, nbuf, tmem_cols;
  int tf32;                   // operands are fp32 in memory, MMA kind::tf32 (K = 8 per instruction)
  int tiles_per_group, groups, col_chunks, nsamples;
  long long total_tiles;
  int ni;                     // MMA issuer warps in use (1 or 2)
  int roll, tiles_per_plane;  // rolling plane window (igemm_roll_kernel): tiles of one plane
  long long tiles_per_chunk, total_steps;
  int seg_len;                // rows of one segment that belong to this tile family
  long long rows_total;       // PLAIN: number of A rows
  // epilogue
  int out_mode, out_fp32, ldc, has_bias, has_stats, stats_per_sample, cout_total;
  int exact_out;              // PCRL_DTYPE_F32X: fp32 results are stored without the tf32 rounding
  int ct_D, ct_H, ct_W;       // coarse dims for the ConvT scatter
  void* out;
  const float* bias;
  double* stats;
};

// the operand kind is a compile-time parameter: a run-time select leaves a predicated-off
// UTC*MMA next to every live one, which costs ~20 % of the tensor-pipe issue rate
template <int TF32>
__device__ __forceinline__ void umma_any(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (TF32) umma_tf32(d, ad, bd, idesc, acc);
  else umma_bf16(d, ad, bd, idesc, acc);
}

// Optional stall accounting (build with -DPCRL_TIMING; tools/bench_layers.py reads it back):
// per CTA [0] MMA-warp loop cycles, waits on [1] a_full [2] b_full [3] acc_empty, [4] slab producer
// on a_empty, [5] filter producer on b_empty, [6] epilogue warp 2 on acc_full, [7] its loop cycles.
#ifdef PCRL_TIMING
__device__ unsigned long long g_timing[1024][8];
#define TWAIT(acc, bar, par) do { long long t0_ = clock64(); mbar_wait(bar, par); acc += clock64() - t0_; } while (0)
#define TDECL(...) long long __VA_ARGS__
#define TNOW() clock64()
#define TSTORE(i, v) do { if (lane == 0) g_timing[blockIdx.x][i] = (unsigned long long)(v); } while (0)
#else
#define TWAIT(acc, bar, par) mbar_wait(bar, par)
#define TDECL(...)
#define TNOW() 0
#define TSTORE(i, v)
#endif

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

struct TileCoord {
  int col0, n, f0, t_local;   // first output column, sample, first flat row, first row inside the segment
};
// step -> tile.  CONV: column chunks outermost, so that CTAs running concurrently share filter tiles in L2 (the
// activations of the layers with more than one chunk fit in L2).  PLAIN GEMM: column chunks INNERMOST -- the CTAs
// that run at the same time then read the same row tiles of A; with chunks outermost the ConvTranspose GEMM of
// up_tr64 (A = 268 MB, eight chunks of 128 columns) re-read A from DRAM once per chunk: 2.2 GB of DRAM reads for
// 0.27 GB of input (ncu, profiles/r02z_ncu_convt.md).  (That kernel is bound by its four epilogue warps, not by
// DRAM: the order removes the traffic, not the time.)
__device__ __forceinline__ TileCoord tile_coord(const IgemmParams& p, long long step) {
  TileCoord c;
  long long chunk, tile;
  if (p.mode == IG_PLAIN) {
    chunk = step % p.col_chunks;
    tile = step / p.col_chunks;
  } else {
    chunk = step / p.tiles_per_chunk;
    tile = step % p.tiles_per_chunk;
  }
  const int t = (int)(tile % p.tiles_per_group);
  long long r = tile / p.tiles_per_group;
  const int g = (int)(r % p.groups);
  r /= p.groups;
  c.n = (int)r;
  c.col0 = (int)chunk * p.nc;
  c.t_local = t * p.m_cta;
  c.f0 = g * p.P * p.PL + c.t_local;
  return c;
}

// Epilogue of one (128-row block, output plane): nc accumulator columns of the 32 rows this warp owns
// (taddr = TMEM address of column 0 for this warp's lanes) -> bias, store at element offset `off`
// (rows of a warp are a row pitch apart), per-channel sum / sum of squares into stat_s.
__device__ __forceinline__ void epilogue_store(const IgemmParams& p, uint32_t taddr, long long off, bool valid,
                                               int co0, float* stat_s, int lane) {
  for (int c = 0; c < p.nc; c += 32) {
    uint32_t v[32];
    tmem_ld32(taddr + c, v);
    tmem_ld_wait();
    float y[32];
#pragma unroll
    for (int i = 0; i < 32; i++) y[i] = __uint_as_float(v[i]);
    if (p.has_bias) {
#pragma unroll
      for (int i = 0; i < 32; i++) y[i] += __ldg(&p.bias[co0 + c + i]);
    }
    if (p.out_mode == OUT_ROWS_T) {
      if (valid) {
        float* o = reinterpret_cast<float*>(p.out) + off;
#pragma unroll
        for (int i = 0; i < 32; i++) o[(size_t)(co0 + c + i) * p.ldc] = y[i];
      }
    } else if (p.out_fp32) {
      if ((p.out_mode == OUT_CONVT || p.out_mode == OUT_UNSHUFFLE) && !p.exact_out) {
        // these outputs are tensor-core operands of the next kernel: store tf32-rounded
#pragma unroll
        for (int i = 0; i < 32; i++) y[i] = rna_tf32(y[i]);
      }
      if (valid) {
        // 256-bit stores: a lane's 32 B fill a whole sector per instruction (the rows of a
        // warp are a row pitch apart, so nothing else coalesces)
        float* o = reinterpret_cast<float*>(p.out) + off + c;
#pragma unroll
        for (int i = 0; i < 4; i++)
          st_global_v8(o + 8 * i, __float_as_uint(y[8 * i]), __float_as_uint(y[8 * i + 1]),
                       __float_as_uint(y[8 * i + 2]), __float_as_uint(y[8 * i + 3]),
                       __float_as_uint(y[8 * i + 4]), __float_as_uint(y[8 * i + 5]),
                       __float_as_uint(y[8 * i + 6]), __float_as_uint(y[8 * i + 7]));
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
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off + c;
#pragma unroll
        for (int i = 0; i < 2; i++)
          st_global_v8(o + 16 * i, pk[8 * i], pk[8 * i + 1], pk[8 * i + 2], pk[8 * i + 3],
                       pk[8 * i + 4], pk[8 * i + 5], pk[8 * i + 6], pk[8 * i + 7]);
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

template <int TF32>
__global__ void __launch_bounds__(256, 1)
igemm_kmajor_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb1,
                    const __grid_constant__ CUtensorMap tb2, const __grid_constant__ CUtensorMap tb3,
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
  uint64_t* acc_full = b_empty + p.sb;     // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
  float* stat_s = (float*)(tmem_slot + 2);  // [2][nc]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.sa; i++) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], (uint32_t)p.ni); }
    for (int i = 0; i < p.sb; i++) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], (uint32_t)p.ni); }
    for (int i = 0; i < 2; i++) { mbar_init(&acc_full[i], (uint32_t)p.ni); mbar_init(&acc_empty[i], 128); }
    fence_barrier_init();
    tma_prefetch_desc(&ta);
    tma_prefetch_desc(&tb1);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  if (p.has_stats)
    for (int i = threadIdx.x; i < 2 * p.nc; i += blockDim.x) stat_s[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const long long first_step = (long long)blockIdx.x, step_stride = (long long)gridDim.x;
  const uint32_t tmem = *tmem_slot;
  const int acc_cols = p.mt * p.P * p.nc;
  const int nslab = (p.mode == IG_CONV) ? p.P + 2 : 1;

  if (warp == 0) {
    // =============================== TMA producer for the activation slabs (whole warp runs the
    // loop so that every operand stays warp-uniform; one elected lane issues)
    int sa = 0, pa = 0;
    TDECL(tw = 0);
    const uint32_t a_tx = (p.mode == IG_CONV) ? (uint32_t)(p.nh_box * p.Wp * p.row_bytes)
                                              : (uint32_t)(p.mt * 128 * p.row_bytes);
    for (long long step = first_step; step < p.total_steps; step += step_stride) {
      const TileCoord tc = tile_coord(p, step);
      int mr_first = 0;
      if (p.mode == IG_CONV) mr_first = floordiv(tc.f0 - p.Wp - 1, p.Wp);
      for (int kb = 0; kb < p.kblocks; kb++) {
        for (int s = 0; s < nslab; s++) {
          const int q = s - 1;
          TWAIT(tw, &a_empty[sa], pa ^ 1);
          uint8_t* dst = a_s + (size_t)sa * p.slab_bytes;
          if (elect_one()) {
            mbar_expect_tx(&a_full[sa], a_tx);
            if (p.mode == IG_CONV) {
              tma_load_4d(dst, &ta, &a_full[sa], kb * p.kc, -1, mr_first + q * p.H1, tc.n);
            } else {
              for (int i = 0; i < p.mt; i++)
                tma_load_2d(dst + (size_t)i * 128 * p.row_bytes, &ta, &a_full[sa], kb * p.kc,
                            tc.f0 + i * 128);
            }
          }
          __syncwarp();
          if (++sa == p.sa) { sa = 0; pa ^= 1; }
        }
      }
    }
    TSTORE(4, tw);
  } else if (warp == 6) {
    // =============================== TMA producer for the filter tiles: its own warp, so weight
    // prefetch runs sb tiles ahead of the MMAs independently of the slab ring
    int sb = 0, pb = 0;
    TDECL(tw = 0);
    for (long long step = first_step; step < p.total_steps; step += step_stride) {
      const TileCoord tc = tile_coord(p, step);
      for (int kb = 0; kb < p.kblocks; kb++) {
        for (int s = 0; s < nslab; s++) {
          const int q = s - 1;
          int cnt = 1, dzr_lo = 0;
          if (p.mode == IG_CONV) {
            const int p_lo = max(0, q - 1), p_hi = min(p.P - 1, q + 1);
            cnt = p_hi - p_lo + 1;
            dzr_lo = p_lo - (q - 1);
          }
          const CUtensorMap* tb = cnt == 1 ? &tb1 : (cnt == 2 ? &tb2 : &tb3);
          const uint32_t b_tx = (uint32_t)(cnt * p.nc * p.row_bytes);
          for (int j = 0; j < p.tpg; j++) {
            TWAIT(tw, &b_empty[sb], pb ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&b_full[sb], b_tx);
              tma_load_3d(b_s + (size_t)sb * p.b_bytes, tb, &b_full[sb], kb * p.kc, tc.col0, j * 3 + dzr_lo);
            }
            __syncwarp();
            if (++sb == p.sb) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
    TSTORE(5, tw);
  } else if (warp == 1 || warp == 7) {
    // =============================== MMA issuers (whole warp loops, one elected lane issues).
    // The issue thread is the critical resource: a tcgen05.mma of N <= 128 retires in 48..69 clk
    // (tools/umma_probe lean), so everything between two MMAs has to be a handful of uniform-
    // register adds.  Descriptors are advanced incrementally (tap -> +1 row / +Wp-2 rows), the
    // k-step loop is straight-line code, and the rare first-touch split lives in a slow path.
    const int iw = (warp == 1) ? 0 : 1;                   // this issuer owns blocks mt = iw, iw + ni, ..
    int sa = 0, pa = 0, sb = 0, pb = 0;
    const uint32_t layout = (p.row_bytes == 128) ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint64_t desc_hi = make_smem_desc(0, 16, 8 * p.row_bytes, layout);
    const bool ks4 = p.row_bytes == 128;                  // 4 (128 B rows) or 2 (64 B rows) k-steps of 32 B
    const uint32_t fmt = TF32 ? 2u : 1u;
    const uint32_t idesc_nc = make_idesc(fmt, 128, (uint32_t)p.nc, 0, 0);
    const uint32_t rb16 = (uint32_t)p.row_bytes >> 4;     // descriptor address units per operand row
    const uint32_t mt16 = 128u * rb16, slab16 = (uint32_t)p.slab_bytes >> 4, bst16 = (uint32_t)p.b_bytes >> 4;
    const uint64_t a_desc0 = desc_hi | (uint64_t)((smem_u32(a_s) >> 4) & 0x3FFF);
    const uint64_t b_desc0 = desc_hi | (uint64_t)((smem_u32(b_s) >> 4) & 0x3FFF);
    const uint32_t d_mt = (uint32_t)(p.P * p.nc);
    const bool conv = p.mode == IG_CONV;
    int it = 0;
    TDECL(twa = 0, twb = 0, twc = 0, tstart = TNOW());
    for (long long step = first_step; step < p.total_steps && iw < p.ni; step += step_stride) {
      const TileCoord tc = tile_coord(p, step);
      const int buf = (p.nbuf == 2) ? (it & 1) : 0;
      const int use = (p.nbuf == 2) ? (it >> 1) : it;      // how often this buffer was used before
      it++;
      TWAIT(twc, &acc_empty[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem + buf * acc_cols;
      // first tap (dy,dx) = (-1,-1): rows before the tile's first row inside the slab
      int tap0_rows = 0;
      if (conv) tap0_rows = tc.f0 - floordiv(tc.f0 - p.Wp - 1, p.Wp) * p.Wp - p.Wp - 1;
      for (int kb = 0; kb < p.kblocks; kb++) {
        for (int s = 0; s < nslab; s++) {
          const int q = s - 1;
          int p_lo = 0, cnt = 1;
          if (conv) {
            p_lo = max(0, q - 1);
            cnt = min(p.P - 1, q + 1) - p_lo + 1;
          }
          // the block of output plane q+1 is touched for the first time by this slab
          const bool has_new = (kb == 0) && (!conv || q + 1 <= p.P - 1);
          const uint32_t idesc_all = make_idesc(fmt, 128, (uint32_t)(cnt * p.nc), 0, 0);
          const uint32_t d0 = acc + (uint32_t)(p_lo * p.nc) + (uint32_t)iw * d_mt;
          TWAIT(twa, &a_full[sa], pa);
          uint64_t ad_tap = a_desc0 + (uint32_t)sa * slab16 + (uint32_t)tap0_rows * rb16 + (uint32_t)iw * mt16;
          int c3 = 0;
          for (int j = 0; j < p.tpg; j++) {
            TWAIT(twb, &b_full[sb], pb);
            tc_fence_after();
            const uint64_t bd = b_desc0 + (uint32_t)sb * bst16;
            if (elect_one()) {
              if (has_new && j == 0) {
                // slow path, once per (tile, slab): old blocks accumulate, the new block (last of
                // the range) is overwritten by its first MMA
                const uint32_t idesc_old = make_idesc(fmt, 128, (uint32_t)((cnt > 1 ? cnt - 1 : 1) * p.nc), 0, 0);
                const uint32_t boff16 = (uint32_t)((cnt - 1) * p.nc) * rb16;
                const int ksteps = ks4 ? 4 : 2;
                for (int mt = iw; mt < p.mt; mt += p.ni) {
                  const uint64_t ad = ad_tap + (uint32_t)(mt - iw) * mt16;
                  const uint32_t d = d0 + (uint32_t)(mt - iw) * d_mt;
                  if (cnt > 1) umma_any<TF32>(d, ad, bd, idesc_old, 1u);
                  umma_any<TF32>(d + (uint32_t)((cnt - 1) * p.nc), ad, bd + boff16, idesc_nc, 0u);
                  for (int ks = 1; ks < ksteps; ks++) umma_any<TF32>(d, ad + 2 * ks, bd + 2 * ks, idesc_all, 1u);
                }
              } else {
                uint64_t ad = ad_tap;
                uint32_t d = d0;
                for (int mt = iw; mt < p.mt; mt += p.ni) {
                  umma_any<TF32>(d, ad, bd, idesc_all, 1u);
                  umma_any<TF32>(d, ad + 2, bd + 2, idesc_all, 1u);
                  if (ks4) {
                    umma_any<TF32>(d, ad + 4, bd + 4, idesc_all, 1u);
                    umma_any<TF32>(d, ad + 6, bd + 6, idesc_all, 1u);
                  }
                  ad += (uint32_t)p.ni * mt16;
                  d += (uint32_t)p.ni * d_mt;
                }
              }
              umma_commit(&b_empty[sb]);
            }
            __syncwarp();
            if (++sb == p.sb) { sb = 0; pb ^= 1; }
            // next tap: dx+1, or wrap to (dy+1, dx=-1)
            if (++c3 == 3) { c3 = 0; ad_tap += (uint32_t)(p.Wp - 2) * rb16; }
            else ad_tap += rb16;
          }
          if (elect_one()) umma_commit(&a_empty[sa]);
          __syncwarp();
          if (++sa == p.sa) { sa = 0; pa ^= 1; }
        }
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
    if (iw == 0) { TSTORE(0, TNOW() - tstart); TSTORE(1, twa); TSTORE(2, twb); TSTORE(3, twc); }
  } else if (warp >= 2 && warp <= 5) {
    // =============================== epilogue warps (warp w owns TMEM lanes 32*(w%4) .. +31)
    const int quad = warp & 3;
    int it = 0;
    int cur_col0 = -1, cur_n = -1;
    const int et = threadIdx.x - 64;  // 0..127
    TDECL(twf = 0, tstart = TNOW());
    for (long long step = first_step; step < p.total_steps; step += step_stride) {
      const TileCoord tc = tile_coord(p, step);
      const int buf = (p.nbuf == 2) ? (it & 1) : 0;
      const int use = (p.nbuf == 2) ? (it >> 1) : it;
      it++;
      if (p.has_stats && (tc.col0 != cur_col0 || (p.stats_per_sample && tc.n != cur_n))) {
        // flush the per-CTA partial statistics of the previous (column chunk, sample)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (cur_col0 >= 0) {
          double* st = p.stats + (p.stats_per_sample ? (size_t)cur_n * p.cout_total * 2 : 0);
          for (int i = et; i < p.nc; i += 128) {
            atomicAdd(&st[(size_t)(cur_col0 + i) * 2 + 0], (double)stat_s[i]);
            atomicAdd(&st[(size_t)(cur_col0 + i) * 2 + 1], (double)stat_s[p.nc + i]);
            stat_s[i] = 0.f;
            stat_s[p.nc + i] = 0.f;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_col0 = tc.col0;
        cur_n = tc.n;
      }
      TWAIT(twf, &acc_full[buf], use & 1);
      tc_fence_after();
      const uint32_t acc = tmem + buf * acc_cols;
      const int t_idx = (p.out_mode == OUT_CONVT) ? tc.col0 / p.cout_total : 0;
      const int co0 = (p.out_mode == OUT_CONVT) ? tc.col0 % p.cout_total : tc.col0;
      for (int mt = 0; mt < p.mt; mt++) {
        const int lrow = mt * 128 + quad * 32 + lane;   // row inside the tile
        for (int pp = 0; pp < p.P; pp++) {
          bool valid;
          long long off;
          if (p.out_mode == OUT_FLAT || p.out_mode == OUT_UNSHUFFLE) {
            const int f = tc.f0 + pp * p.PL + lrow;
            const int mr = f / p.Wp, wq = f - mr * p.Wp;
            const int hp = mr % p.H1;
            valid = (wq >= 1) && (mr < p.MR) && (hp >= 1) && (tc.t_local + lrow < p.seg_len);
            if (p.out_mode == OUT_FLAT) {
              off = (((long long)tc.n * p.MR + mr) * p.W + (wq - 1)) * p.ldc + co0;
            } else {
              // fine voxel (d2, h2, w2) -> coarse-major [coarse H-padded row][tap][C] (the layout
              // the ConvTranspose gradient GEMMs consume); ct_* are the coarse dims
              const int d2 = mr / p.H1, h2 = hp - 1, w2 = wq - 1;
              const int t = ((d2 & 1) << 2) | ((h2 & 1) << 1) | (w2 & 1);
              const long long crow = (((long long)tc.n * p.ct_D + (d2 >> 1)) * (p.ct_H + 1) + (h2 >> 1) + 1) * p.ct_W + (w2 >> 1);
              off = (crow * 8 + t) * p.ldc + co0;
            }
          } else if (p.out_mode == OUT_ROWS) {
            const long long r = (long long)tc.f0 + lrow;
            valid = r < p.rows_total;
            off = r * p.ldc + co0;
          } else if (p.out_mode == OUT_ROWS_T) {
            // transposed fp32 store out[col][row]: lanes hold consecutive rows -> coalesced
            const long long r = (long long)tc.f0 + lrow;
            valid = r < p.rows_total;
            off = r;
          } else {
            // coarse H-padded row r = ((n*D + d)*(H+1) + h')*W + w  ->  fine voxel (2d+i, 2h+j, 2w+k)
            const long long r = (long long)tc.f0 + lrow;
            valid = r < p.rows_total;
            long long qq = r;
            const int w = (int)(qq % p.ct_W); qq /= p.ct_W;
            const int hp = (int)(qq % (p.ct_H + 1)); qq /= (p.ct_H + 1);
            const int d = (int)(qq % p.ct_D);
            const long long nn = qq / p.ct_D;
            valid = valid && hp >= 1;
            const int i = t_idx >> 2, j = (t_idx >> 1) & 1, k = t_idx & 1;
            const long long fd = 2 * d + i, fh = 2 * (hp - 1) + j + 1, fw = 2 * w + k;
            off = (((nn * (2 * p.ct_D) + fd) * (2 * p.ct_H + 1) + fh) * (2 * p.ct_W) + fw) * p.ldc + co0;
          }
          epilogue_store(p, acc + ((uint32_t)(quad * 32) << 16) + (mt * p.P + pp) * p.nc, off, valid, co0, stat_s, lane);
        }
      }
      // accumulator buffer drained: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
    if (p.has_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (cur_col0 >= 0) {
        double* st = p.stats + (p.stats_per_sample ? (size_t)cur_n * p.cout_total * 2 : 0);
        for (int i = et; i < p.nc; i += 128) {
          atomicAdd(&st[(size_t)(cur_col0 + i) * 2 + 0], (double)stat_s[i]);
          atomicAdd(&st[(size_t)(cur_col0 + i) * 2 + 1], (double)stat_s[p.nc + i]);
        }
      }
    }
    if (warp == 2) { TSTORE(6, twf); TSTORE(7, TNOW() - tstart); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------
// ROLLING PLANE WINDOW variant of the convolution kernel, for the layers whose column chunk is narrow
// (nc <= 64: every full-resolution 64-channel layer, i.e. most of the FLOPs).
//
// Why: a tcgen05.mma occupies the pipe for >= ~68 clk whatever its width N <= 128 (tools/umma_probe), so an
// N = 64 MMA runs at less than half rate.  The kernel above therefore stacks the <= 3 output planes that one
// input plane feeds along N, but with tiles of P = 2 output planes the first and last input plane of every
// tile feed ONE plane each (N = 64): per tile 74 + 68 + 68 + 74 clk for 6 x 34 clk of math = 72 % -- the
// measured tensor-pipe activity of those layers (68-71 %, profiles/r01c_ncu_tensor_kernels.md).
// Here a CTA owns a 256-row window of the plane and walks through ALL D planes: input plane q is loaded once
// (not (P+2)/P times) and feeds output planes q-1, q, q+1 in ONE N = 3*nc MMA per tap; the accumulators of
// four consecutive output planes live in a TMEM ring (4 slots x nc columns per 128-row block), plane q-1 is
// handed to the epilogue as soon as input plane q is done, and its slot is reused four planes later -- the
// epilogue of plane p overlaps the MMAs of planes p+1.. without a second accumulator buffer.
// When the three active slots wrap around the ring (2 of 4 positions) the MMA is split in two.
// Requirements: D % 4 == 0, nc <= 64, a plane of at least 512 flat rows.
template <int TF32>
__global__ void __launch_bounds__(256, 1)
igemm_roll_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb2,
                  const __grid_constant__ CUtensorMap tb3, const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;
  uint8_t* b_s = a_s + (size_t)p.sa * p.slab_bytes;
  uint64_t* bars = (uint64_t*)(b_s + (size_t)p.sb * p.b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.sa;
  uint64_t* b_full = a_empty + p.sa;
  uint64_t* b_empty = b_full + p.sb;
  uint64_t* acc_full = b_empty + p.sb;     // [4]: one per ring slot
  uint64_t* acc_empty = acc_full + 4;      // [4]
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 4);
  float* stat_s = (float*)(tmem_slot + 2);  // [2][nc]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.sa; i++) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], (uint32_t)p.ni); }
    for (int i = 0; i < p.sb; i++) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], (uint32_t)p.ni); }
    for (int i = 0; i < 4; i++) { mbar_init(&acc_full[i], (uint32_t)p.ni); mbar_init(&acc_empty[i], 128); }
    fence_barrier_init();
    tma_prefetch_desc(&ta);
    tma_prefetch_desc(&tb3);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  if (p.has_stats)
    for (int i = threadIdx.x; i < 2 * p.nc; i += blockDim.x) stat_s[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const long long first_step = (long long)blockIdx.x, step_stride = (long long)gridDim.x;
  const uint32_t tmem = *tmem_slot;
  const int D = p.D, upt = p.D >> 2;                // uses of every ring slot per tile
  const int blk_cols = 4 * p.nc;                    // accumulator columns of one 128-row block (4 slots)
  const long long tiles_per_chunk = (long long)p.tiles_per_plane * p.nsamples;

  if (warp == 0) {
    // =============================== activation slabs: input plane q of this window, once per k-block
    int sa = 0, pa = 0;
    const uint32_t a_tx = (uint32_t)(p.nh_box * p.Wp * p.row_bytes);
    for (long long step = first_step; step < p.total_steps; step += step_stride) {
      const long long tile = step % tiles_per_chunk;
      const int t_local = (int)(tile % p.tiles_per_plane) * p.m_cta, n = (int)(tile / p.tiles_per_plane);
      const int mr_first = floordiv(t_local - p.Wp - 1, p.Wp);
      for (int q = 0; q < D; q++)
        for (int kb = 0; kb < p.kblocks; kb++) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&a_full[sa], a_tx);
            tma_load_4d(a_s + (size_t)sa * p.slab_bytes, &ta, &a_full[sa], kb * p.kc, -1, mr_first + q * p.H1, n);
          }
          __syncwarp();
          if (++sa == p.sa) { sa = 0; pa ^= 1; }
        }
    }
  } else if (warp == 6) {
    // =============================== filter tiles: per (input plane, k-block, (dy,dx) tap) the 2 or 3 dz taps
    int sb = 0, pb = 0;
    for (long long step = first_step; step < p.total_steps; step += step_stride) {
      const int col0 = (int)(step / tiles_per_chunk) * p.nc;
      for (int q = 0; q < D; q++) {
        const int lo = max(0, q - 1), hi = min(D - 1, q + 1), cnt = hi - lo + 1, dzr_lo = lo - (q - 1);
        const CUtensorMap* tb = cnt == 2 ? &tb2 : &tb3;
        const uint32_t b_tx = (uint32_t)(cnt * p.nc * p.row_bytes);
        for (int kb = 0; kb < p.kblocks; kb++)
          for (int j = 0; j < 9; j++) {
            mbar_wait(&b_empty[sb], pb ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&b_full[sb], b_tx);
              tma_load_3d(b_s + (size_t)sb * p.b_bytes, tb, &b_full[sb], kb * p.kc, col0, j * 3 + dzr_lo);
            }
            __syncwarp();
            if (++sb == p.sb) { sb = 0; pb ^= 1; }
          }
      }
    }
  } else if (warp == 1 || warp == 7) {
    // =============================== MMA issuers: issuer iw owns 128-row block iw of the window
    const int iw = (warp == 1) ? 0 : 1;
    int sa = 0, pa = 0, sb = 0, pb = 0;
    const uint32_t layout = (p.row_bytes == 128) ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint64_t desc_hi = make_smem_desc(0, 16, 8 * p.row_bytes, layout);
    const int ksteps = (p.row_bytes == 128) ? 4 : 2;
    const uint32_t fmt = TF32 ? 2u : 1u;
    const uint32_t idesc1 = make_idesc(fmt, 128, (uint32_t)p.nc, 0, 0);
    const uint32_t idesc2 = make_idesc(fmt, 128, (uint32_t)(2 * p.nc), 0, 0);
    const uint32_t idesc3 = make_idesc(fmt, 128, (uint32_t)(3 * p.nc), 0, 0);
    const uint32_t rb16 = (uint32_t)p.row_bytes >> 4;
    const uint32_t mt16 = 128u * rb16, slab16 = (uint32_t)p.slab_bytes >> 4, bst16 = (uint32_t)p.b_bytes >> 4;
    const uint32_t nc16 = (uint32_t)p.nc * rb16;           // descriptor units of one plane's filter rows
    const uint64_t a_desc0 = desc_hi | (uint64_t)((smem_u32(a_s) >> 4) & 0x3FFF);
    const uint64_t b_desc0 = desc_hi | (uint64_t)((smem_u32(b_s) >> 4) & 0x3FFF);
    const uint32_t dblk = tmem + (uint32_t)(iw * blk_cols);
    int it = 0;
    for (long long step = first_step; step < p.total_steps && iw < p.ni; step += step_stride, it++) {
      const long long tile = step % tiles_per_chunk;
      const int t_local = (int)(tile % p.tiles_per_plane) * p.m_cta;
      const int tap0_rows = t_local - floordiv(t_local - p.Wp - 1, p.Wp) * p.Wp - p.Wp - 1;
      for (int q = 0; q < D; q++) {
        const int lo = max(0, q - 1), hi = min(D - 1, q + 1), cnt = hi - lo + 1;
        // output planes touched for the first time by this input plane: their ring slot must have been drained
        const bool has_new = (q == 0) || (q + 1 <= D - 1);
        if (q == 0) {
          mbar_wait(&acc_empty[0], ((it * upt) & 1) ^ 1);
          mbar_wait(&acc_empty[1], ((it * upt) & 1) ^ 1);
        } else if (q + 1 <= D - 1) {
          mbar_wait(&acc_empty[(q + 1) & 3], ((it * upt + ((q + 1) >> 2)) & 1) ^ 1);
        }
        tc_fence_after();
        const int s_lo = lo & 3;
        const int n1 = min(cnt, 4 - s_lo);                 // planes before the ring wraps
        const uint32_t id_a = n1 == 1 ? idesc1 : (n1 == 2 ? idesc2 : idesc3);
        const uint32_t id_b = (cnt - n1) == 1 ? idesc1 : idesc2;
        const uint32_t d_a = dblk + (uint32_t)(s_lo * p.nc);
        for (int kb = 0; kb < p.kblocks; kb++) {
          mbar_wait(&a_full[sa], pa);
          uint64_t ad = a_desc0 + (uint32_t)sa * slab16 + (uint32_t)tap0_rows * rb16 + (uint32_t)iw * mt16;
          int c3 = 0;
          for (int j = 0; j < 9; j++) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint64_t bd = b_desc0 + (uint32_t)sb * bst16;
            if (elect_one()) {
              if (has_new && kb == 0 && j == 0) {
                // first tap of an input plane that opens new output planes: plane by plane, the new ones
                // are overwritten by their first k-step
                for (int r = 0; r < cnt; r++) {
                  const int pl = lo + r;
                  const bool is_new = (q == 0) || (pl == q + 1);
                  const uint32_t d = dblk + (uint32_t)((pl & 3) * p.nc);
                  for (int ks = 0; ks < ksteps; ks++)
                    umma_any<TF32>(d, ad + 2 * ks, bd + (uint32_t)r * nc16 + 2 * ks, idesc1, (ks > 0 || !is_new) ? 1u : 0u);
                }
              } else {
                umma_any<TF32>(d_a, ad, bd, id_a, 1u);
                umma_any<TF32>(d_a, ad + 2, bd + 2, id_a, 1u);
                if (ksteps == 4) {
                  umma_any<TF32>(d_a, ad + 4, bd + 4, id_a, 1u);
                  umma_any<TF32>(d_a, ad + 6, bd + 6, id_a, 1u);
                }
                if (n1 < cnt) {                            // wrapped part: ring slots 0..
                  const uint64_t bd2 = bd + (uint32_t)n1 * nc16;
                  umma_any<TF32>(dblk, ad, bd2, id_b, 1u);
                  umma_any<TF32>(dblk, ad + 2, bd2 + 2, id_b, 1u);
                  if (ksteps == 4) {
                    umma_any<TF32>(dblk, ad + 4, bd2 + 4, id_b, 1u);
                    umma_any<TF32>(dblk, ad + 6, bd2 + 6, id_b, 1u);
                  }
                }
              }
              umma_commit(&b_empty[sb]);
            }
            __syncwarp();
            if (++sb == p.sb) { sb = 0; pb ^= 1; }
            if (++c3 == 3) { c3 = 0; ad += (uint32_t)(p.Wp - 2) * rb16; }
            else ad += rb16;
          }
          if (elect_one()) umma_commit(&a_empty[sa]);
          __syncwarp();
          if (++sa == p.sa) { sa = 0; pa ^= 1; }
        }
        // input plane q done: output plane q-1 is complete (and plane D-1 after the last input plane)
        if (elect_one()) {
          if (q >= 1) umma_commit(&acc_full[(q - 1) & 3]);
          if (q == D - 1) umma_commit(&acc_full[(D - 1) & 3]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 2 && warp <= 5) {
    // =============================== epilogue warps (warp w owns TMEM lanes 32*(w%4) .. +31)
    const int quad = warp & 3;
    int it = 0;
    int cur_col0 = -1, cur_n = -1;
    const int et = threadIdx.x - 64;
    for (long long step = first_step; step < p.total_steps; step += step_stride, it++) {
      const long long tile = step % tiles_per_chunk;
      const int col0 = (int)(step / tiles_per_chunk) * p.nc;
      const int t_local = (int)(tile % p.tiles_per_plane) * p.m_cta, n = (int)(tile / p.tiles_per_plane);
      if (p.has_stats && (col0 != cur_col0 || (p.stats_per_sample && n != cur_n))) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (cur_col0 >= 0) {
          double* st = p.stats + (p.stats_per_sample ? (size_t)cur_n * p.cout_total * 2 : 0);
          for (int i = et; i < p.nc; i += 128) {
            atomicAdd(&st[(size_t)(cur_col0 + i) * 2 + 0], (double)stat_s[i]);
            atomicAdd(&st[(size_t)(cur_col0 + i) * 2 + 1], (double)stat_s[p.nc + i]);
            stat_s[i] = 0.f;
            stat_s[p.nc + i] = 0.f;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_col0 = col0;
        cur_n = n;
      }
      for (int pl = 0; pl < D; pl++) {
        const int slot = pl & 3;
        mbar_wait(&acc_full[slot], (it * upt + (pl >> 2)) & 1);
        tc_fence_after();
        for (int mt = 0; mt < p.mt; mt++) {
          const int lrow = mt * 128 + quad * 32 + lane;
          const int f = pl * p.PL + t_local + lrow;
          const int mr = f / p.Wp, wq = f - mr * p.Wp;
          const int hp = mr % p.H1;
          const bool valid = (wq >= 1) && (mr < p.MR) && (hp >= 1) && (t_local + lrow < p.PL);
          long long off;
          if (p.out_mode == OUT_FLAT) {
            off = (((long long)n * p.MR + mr) * p.W + (wq - 1)) * p.ldc + col0;
          } else {   // OUT_UNSHUFFLE: coarse-major [coarse H-padded row][tap][C]
            const int d2 = mr / p.H1, h2 = hp - 1, w2 = wq - 1;
            const int t = ((d2 & 1) << 2) | ((h2 & 1) << 1) | (w2 & 1);
            const long long crow = (((long long)n * p.ct_D + (d2 >> 1)) * (p.ct_H + 1) + (h2 >> 1) + 1) * p.ct_W + (w2 >> 1);
            off = (crow * 8 + t) * p.ldc + col0;
          }
          epilogue_store(p, tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * blk_cols + slot * p.nc), off, valid,
                         col0, stat_s, lane);
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[slot]);
      }
    }
    if (p.has_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (cur_col0 >= 0) {
        double* st = p.stats + (p.stats_per_sample ? (size_t)cur_n * p.cout_total * 2 : 0);
        for (int i = et; i < p.nc; i += 128) {
          atomicAdd(&st[(size_t)(cur_col0 + i) * 2 + 0], (double)stat_s[i]);
          atomicAdd(&st[(size_t)(cur_col0 + i) * 2 + 1], (double)stat_s[p.nc + i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------ host side
static int pow2_cols(int c) {
  int t = 32;
  while (t < c) t <<= 1;
  return t;
}

static inline CUtensorMapDataType tma_dtype(int tf32) {
  return tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}
static inline CUtensorMapSwizzle tma_swizzle(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
}

struct IgemmLaunch {
  IgemmParams p;
  CUtensorMap ta, tb[3];
};

static int launch_roll(IgemmLaunch& L, cudaStream_t stream) {
  IgemmParams& p = L.p;
  p.tmem_cols = pow2_cols(p.mt * 4 * p.nc);
  p.b_bytes = ((3 * p.nc * p.row_bytes + 1023) / 1024) * 1024;
  const size_t budget = 212 * 1024;
  p.sa = 2;
  p.sb = 2;
  if ((size_t)p.sa * p.slab_bytes + (size_t)p.sb * p.b_bytes > budget)
    return fail(PCRL_ERR_ARG, "igemm (rolling): tile does not fit shared memory (slab %d B, b %d B)", p.slab_bytes, p.b_bytes);
  while ((size_t)p.sa * p.slab_bytes + (size_t)(p.sb + 1) * p.b_bytes <= budget && p.sb < 6) p.sb++;
  if ((size_t)(p.sa + 1) * p.slab_bytes + (size_t)p.sb * p.b_bytes <= budget) p.sa++;
  while ((size_t)p.sa * p.slab_bytes + (size_t)(p.sb + 1) * p.b_bytes <= budget && p.sb < 10) p.sb++;
  const size_t smem = (size_t)p.sa * p.slab_bytes + (size_t)p.sb * p.b_bytes +
                      (2 * p.sa + 2 * p.sb + 8) * 8 + 16 + 2 * p.nc * 4 + 1024;
  static bool configured = false;
  if (!configured) {
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(igemm_roll_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(igemm_roll_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  p.ni = 2;
  p.total_steps = (long long)p.tiles_per_plane * p.nsamples * p.col_chunks;
  long long grid = num_sms();
  if (p.total_steps < grid) grid = p.total_steps;
  if (p.tf32) igemm_roll_kernel<1><<<(unsigned)grid, 256, smem, stream>>>(L.ta, L.tb[1], L.tb[2], p);
  else igemm_roll_kernel<0><<<(unsigned)grid, 256, smem, stream>>>(L.ta, L.tb[1], L.tb[2], p);
  PCRL_CHECK_CUDA(cudaGetLastError());
  return PCRL_OK;
}

static int finish_and_launch(IgemmLaunch& L, cudaStream_t stream) {
  IgemmParams& p = L.p;
  const int acc_cols = p.mt * p.P * p.nc;
  if (acc_cols > 512) return fail(PCRL_ERR_ARG, "igemm: mt*P*nc=%d exceeds TMEM", acc_cols);
  p.nbuf = (2 * acc_cols <= 512) ? 2 : 1;
  p.tmem_cols = pow2_cols(p.nbuf * acc_cols);
  const int cnt_max = (p.mode == IG_CONV) ? (p.P >= 3 ? 3 : (p.P == 2 ? 2 : 1)) : 1;
  p.b_bytes = ((cnt_max * p.nc * p.row_bytes + 1023) / 1024) * 1024;
  // stage counts: a slab lasts nine taps, so two (three when cheap) are enough; everything else goes
  // to the filter ring, whose depth is what hides the L2 latency of the per-tap weight tiles
  const size_t budget = 212 * 1024;
  p.sa = 2;
  p.sb = 2;
  if ((size_t)p.sa * p.slab_bytes + (size_t)p.sb * p.b_bytes > budget)
    return fail(PCRL_ERR_ARG, "igemm: tile does not fit shared memory (slab %d B, b %d B)",
                p.slab_bytes, p.b_bytes);
  if (p.mode == IG_PLAIN) {
    while ((size_t)(p.sa + 1) * (p.slab_bytes + p.b_bytes) <= budget && p.sa < 6) { p.sa++; p.sb++; }
  } else {
    while ((size_t)p.sa * p.slab_bytes + (size_t)(p.sb + 1) * p.b_bytes <= budget && p.sb < 6) p.sb++;
    if ((size_t)(p.sa + 1) * p.slab_bytes + (size_t)p.sb * p.b_bytes <= budget) p.sa++;
    while ((size_t)p.sa * p.slab_bytes + (size_t)(p.sb + 1) * p.b_bytes <= budget && p.sb < 10) p.sb++;
  }
  const size_t smem = (size_t)p.sa * p.slab_bytes + (size_t)p.sb * p.b_bytes +
                      (2 * p.sa + 2 * p.sb + 4) * 8 + 16 + 2 * p.nc * 4 + 1024;
  static bool configured = false;
  if (!configured) {
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(igemm_kmajor_kernel<0>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PCRL_CHECK_CUDA(cudaFuncSetAttribute(igemm_kmajor_kernel<1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  p.ni = p.mt >= 2 ? 2 : 1;
  p.tiles_per_chunk = (long long)p.tiles_per_group * p.groups * p.nsamples;
  p.total_steps = p.tiles_per_chunk * p.col_chunks;
  long long grid = num_sms();
  if (p.total_steps < grid) grid = p.total_steps;
  if (p.tf32)
    igemm_kmajor_kernel<1><<<(unsigned)grid, 256, smem, stream>>>(L.ta, L.tb[0], L.tb[1], L.tb[2], p);
  else
    igemm_kmajor_kernel<0><<<(unsigned)grid, 256, smem, stream>>>(L.ta, L.tb[0], L.tb[1], L.tb[2], p);
  PCRL_CHECK_CUDA(cudaGetLastError());
  return PCRL_OK;
}

// B operand maps: packed weights viewed as (K, cols, taps) bf16, boxes (kc, nc, cnt), cnt = 1..3
static int make_b_maps(IgemmLaunch& L, const void* w, int K, int cols, int taps) {
  const uint64_t elt = L.p.tf32 ? 4 : 2;
  for (int cnt = 1; cnt <= 3; cnt++) {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)cols, (uint64_t)taps};
    uint64_t str[2] = {(uint64_t)K * elt, (uint64_t)K * cols * elt};
    uint32_t box[3] = {(uint32_t)L.p.kc, (uint32_t)L.p.nc, (uint32_t)(cnt <= taps ? cnt : 1)};
    int rc = encode_map(&L.tb[cnt - 1], tma_dtype(L.p.tf32), 3, w, dims, str, box, tma_swizzle(L.p.row_bytes));
    if (rc) return rc;
  }
  return PCRL_OK;
}

// 3x3x3 convolution (forward or data gradient) on H-padded NDHWC bf16 activations.
//   x: [N][D][H+1][W][Cin], w: [9 (ky,kx)][3 (kz = 2,1,0)][Cout][Cin] bf16 (pcrl_pack_conv3_weights),
//   y: [N][D][H+1][W][Cout] (bf16, or fp32)
//   stats (optional): [Cout][2] (or [N][Cout][2]) fp64, ACCUMULATED (caller zeroes)
int conv3d_k3_igemm(const void* x, const void* w, void* y, double* stats, int stats_per_sample,
                    int out_fp32, int N, int D, int H, int W, int Cin, int Cout,
                    cudaStream_t stream, int unshuffle, int dtype) {
  const int tf32 = dtype != PCRL_DTYPE_BF16;
  if (tf32) {
    PCRL_REQUIRE(Cin % 32 == 0, "conv3d_k3 (fp32): Cin=%d must be a multiple of 32", Cin);
    out_fp32 = 1;
  } else {
    PCRL_REQUIRE(Cin == 32 || Cin % 64 == 0, "conv3d_k3: Cin=%d must be 32 or a multiple of 64", Cin);
  }
  PCRL_REQUIRE(Cout % 32 == 0, "conv3d_k3: Cout=%d must be a multiple of 32", Cout);
  PCRL_REQUIRE(W + 1 <= 256 && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_k3: bad dims");
  IgemmLaunch L;
  memset(&L.p, 0, sizeof(L.p));
  IgemmParams& p = L.p;
  p.mode = IG_CONV;
  p.W = W; p.Wp = W + 1; p.H1 = H + 1; p.D = D; p.MR = D * (H + 1); p.PL = p.H1 * p.Wp;
  p.tf32 = tf32;
  const int elt = tf32 ? 4 : 2;
  p.kc = tf32 ? 32 : ((Cin == 32) ? 32 : 64);
  p.row_bytes = p.kc * elt;
  p.kblocks = Cin / p.kc;
  p.tpg = 9;
  p.nc = (Cout % 128 == 0) ? 128 : (Cout % 64 == 0 ? 64 : 32);
  const long long flat = (long long)p.MR * p.Wp;
  // plane stacking: only where N = nc alone would starve the tensor pipe (nc <= 64) and the
  // planes are large enough for per-plane tiles
  p.P = 1;
  if (p.nc <= 64 && p.PL >= 512) {
    p.P = (p.nc == 64) ? 2 : 4;
    while (p.P > 1 && D % p.P) p.P >>= 1;
  }
  // rolling plane window (igemm_roll_kernel) for the narrow-column layers with large planes
  p.roll = (p.nc <= 64 && p.PL >= 512 && D >= 4 && D % 4 == 0 && !getenv("PCRL_IGEMM_NOROLL")) ? 1 : 0;
  p.mt = (flat >= 256) ? 2 : 1;
  {  // tuning overrides (experiments only)
    const char* e = getenv("PCRL_IGEMM_MT");
    if (e && atoi(e) > 0 && flat >= 128LL * atoi(e)) p.mt = atoi(e);
    e = getenv("PCRL_IGEMM_P");
    if (e && atoi(e) > 0 && p.P > 1) { p.P = atoi(e); while (p.P > 1 && D % p.P) p.P >>= 1; }
    while (p.mt > 1 && p.mt * p.P * p.nc > 512) p.mt >>= 1;
  }
  if (p.roll) { p.mt = 2; p.P = 4; }
  p.m_cta = p.mt * 128;
  if (p.roll) {
    p.seg_len = p.PL;
    p.groups = 1;
    p.tiles_per_plane = (p.PL + p.m_cta - 1) / p.m_cta;
  } else if (p.P > 1) {
    p.seg_len = p.PL;
    p.groups = D / p.P;
  } else {
    p.seg_len = (int)flat;
    p.groups = 1;
  }
  p.tiles_per_group = (p.seg_len + p.m_cta - 1) / p.m_cta;
  p.col_chunks = Cout / p.nc;
  p.nsamples = N;
  p.nh_box = (p.m_cta + 3 * p.Wp + 1 + p.Wp - 1) / p.Wp;
  if (p.nh_box > 256) return fail(PCRL_ERR_UNSUPPORTED, "conv3d_k3: W=%d too small for box", W);
  p.slab_bytes = ((p.nh_box * p.Wp * p.row_bytes + 1023) / 1024) * 1024;
  p.out_mode = unshuffle ? OUT_UNSHUFFLE : OUT_FLAT; p.out_fp32 = out_fp32; p.ldc = Cout; p.cout_total = Cout;
  if (unshuffle) {
    PCRL_REQUIRE(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "conv3d_k3: the coarse-major (unshuffled) output needs even dims");
    p.ct_D = D / 2; p.ct_H = H / 2; p.ct_W = W / 2;
  }
  p.has_stats = stats != nullptr; p.stats_per_sample = stats_per_sample;
  p.out = y; p.stats = stats; p.exact_out = dtype == PCRL_DTYPE_F32X;
  uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)p.MR, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)Cin * elt, (uint64_t)W * Cin * elt, (uint64_t)p.MR * W * Cin * elt};
  uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)p.Wp, (uint32_t)p.nh_box, 1};
  int rc = encode_map(&L.ta, tma_dtype(tf32), 4, x, dims, str, box, tma_swizzle(p.row_bytes));
  if (rc) return rc;
  rc = make_b_maps(L, w, Cin, Cout, 27);
  if (rc) return rc;
  if (p.roll) return launch_roll(L, stream);
  return finish_and_launch(L, stream);
}

// Plain GEMM  C[rows][cols] = A[rows][K] * B[cols][K]^T (+ bias[col]); A, B bf16 row-major.
// out_mode OUT_ROWS stores C row-major with leading dimension ldc; OUT_CONVT scatters the
// (tap, cout) columns of a ConvTranspose3d(k=2,s=2) to the fine H-padded NDHWC tensor.
int gemm_nt_igemm(const void* a, const void* b, void* c, const float* bias, long long rows, int K,
                  int cols, int ldc, int out_fp32, int out_mode, int ct_D, int ct_H, int ct_W,
                  int ct_cout, cudaStream_t stream, int dtype, double* stats) {
  const int tf32 = dtype != PCRL_DTYPE_BF16;
  if (tf32) {
    PCRL_REQUIRE(K % 32 == 0, "gemm_nt (fp32): K=%d must be a multiple of 32", K);
    out_fp32 = 1;
  } else {
    PCRL_REQUIRE(K % 64 == 0 || K == 32, "gemm_nt: K=%d must be 32 or a multiple of 64", K);
  }
  PCRL_REQUIRE(cols % 32 == 0, "gemm_nt: cols=%d must be a multiple of 32", cols);
  PCRL_REQUIRE(rows < (1LL << 31), "gemm_nt: too many rows");
  IgemmLaunch L;
  memset(&L.p, 0, sizeof(L.p));
  IgemmParams& p = L.p;
  p.mode = IG_PLAIN;
  p.tf32 = tf32;
  const int elt = tf32 ? 4 : 2;
  p.kc = tf32 ? 32 : ((K == 32) ? 32 : 64); p.row_bytes = p.kc * elt; p.kblocks = K / p.kc; p.tpg = 1; p.P = 1;
  p.nc = (cols % 128 == 0) ? 128 : (cols % 64 == 0 ? 64 : 32);
  if (out_mode == OUT_CONVT) {
    PCRL_REQUIRE(ct_cout % p.nc == 0, "convT: Cout=%d must be a multiple of %d", ct_cout, p.nc);
  }
  p.mt = rows >= 256 ? 2 : 1;
  p.m_cta = p.mt * 128;
  p.slab_bytes = p.m_cta * p.row_bytes;
  p.rows_total = rows;
  p.seg_len = 0; p.PL = 0; p.Wp = 1;
  p.tiles_per_group = (int)((rows + p.m_cta - 1) / p.m_cta);
  p.groups = 1; p.nsamples = 1; p.col_chunks = cols / p.nc;
  p.out_mode = out_mode; p.out_fp32 = out_fp32; p.ldc = ldc;
  p.cout_total = (out_mode == OUT_CONVT) ? ct_cout : cols;
  p.ct_D = ct_D; p.ct_H = ct_H; p.ct_W = ct_W;
  p.has_bias = bias != nullptr; p.bias = bias; p.out = c; p.exact_out = dtype == PCRL_DTYPE_F32X;
  p.has_stats = stats != nullptr; p.stats = stats; p.stats_per_sample = 0;   // column sums / sums of squares
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)K * elt};
  uint32_t box[2] = {(uint32_t)p.kc, 128};
  int rc = encode_map(&L.ta, tma_dtype(tf32), 2, a, dims, str, box, tma_swizzle(p.row_bytes));
  if (rc) return rc;
  rc = make_b_maps(L, b, K, cols, 1);
  if (rc) return rc;
  return finish_and_launch(L, stream);
}

}  // namespace pcrl

#ifdef PCRL_TIMING
extern "C" int pcrl_debug_timing(unsigned long long* host, int nblocks) {
  return (int)cudaMemcpyFromSymbol(host, pcrl::g_timing, (size_t)nblocks * 8 * sizeof(unsigned long long));
}
#endif
