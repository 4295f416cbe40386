// Host-side helpers shared by the translation units of libpcrl_b200.so: error reporting,
// TMA tensor-map encoding (driver entry point fetched at run time, no -lcuda needed) and the
// activation-layout contract.
//
// ACTIVATION LAYOUT ("H-padded NDHWC"): a tensor with logical shape (N, C, D, H, W) is stored as
//   [N][D][H+1][W][C]   (C fastest), bf16 or fp32,
// where row h' = 0 of every (n, d) plane is all zeros and voxel (d, h, w) lives at row h' = h+1.
// Consecutive planes of one sample are contiguous, so (d, h') merge into one "merged row" index
// mr = d*(H+1) + h' in [0, D*(H+1)).  The 3x3x3 kernels address a sample through the flat index
//   f = mr*(W+1) + (w+1)
// (one virtual zero column per row, produced by TMA out-of-bounds fill, never stored), in which a
// filter tap (dz,dy,dx) is the constant shift  dz*(H+1)*(W+1) + dy*(W+1) + dx.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/pcrl_b200.h"

namespace pcrl {

// error codes: include/pcrl_b200.h

char* last_error_buf();  // defined in api.cu (thread-local, 512 bytes)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define PCRL_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return ::pcrl::fail(PCRL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,           \
                          cudaGetErrorString(e__), __FILE__, __LINE__);                    \
  } while (0)

#define PCRL_CHECK_LAUNCH() PCRL_CHECK_CUDA(cudaGetLastError())

#define PCRL_REQUIRE(cond, ...)                                      \
  do {                                                               \
    if (!(cond)) return ::pcrl::fail(PCRL_ERR_ARG, __VA_ARGS__); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled();  // defined in api.cu

// Encodes a tiled tensor map. dims/box are innermost-first; strides_bytes has rank-1 entries
// (stride of dim 1..rank-1). Returns 0 or a negative error code.
inline int encode_map(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* ptr,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(PCRL_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; i++) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; i++) s[i] = strides_bytes[i];
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(ptr), d, s, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PCRL_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r,
                rank);
  return PCRL_OK;
}

// Round-to-nearest fp32 -> tf32 (10-bit mantissa).  The tensor core truncates fp32 operands to
// tf32; storing operands already rounded makes that truncation exact, halves the operand error
// and removes its bias.
#ifdef __CUDACC__
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
#endif

inline int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

}  // namespace pcrl
