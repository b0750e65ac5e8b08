// Shared helpers for libfrx_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>

#include "../../include/frx.h"

namespace frx {

void set_error(const char* fmt, ...);

#define FRX_CHECK_ARG(cond, ...)                    \
  do {                                              \
    if (!(cond)) {                                  \
      frx::set_error(__VA_ARGS__);                  \
      return FRX_E_ARG;                             \
    }                                               \
  } while (0)

#define FRX_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      frx::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return FRX_E_CUDA;                                                                 \
    }                                                                                    \
  } while (0)

#define FRX_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      frx::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return FRX_E_CUDA;                                                                 \
    }                                                                                    \
  } while (0)

int num_sms();

// score.cu: out[m, n] = scale * sum_k A[m, k] * B[n, k], fp32 operands read as tf32, fp32 accumulation (tcgen05)
// ksplit_ws (optional): scratch for K-split partial tiles when the output covers only a few tiles
int dense_tf32_scaled(const float* a, int64_t ld_a, const float* b, int64_t ld_b, int m, int64_t n, int k, float* out,
                      int64_t ld_out, float scale, void* stream, void* ksplit_ws = nullptr, size_t ksplit_ws_bytes = 0);

// gemm3x.cu: fp32-grade GEMM on the tf32 tensor cores, 3xTF32 operand split inside the kernel (no split / transposed
// copies in HBM).  C[m, n] = act(alpha * (sum_k A(m,k) B(n,k)) * col_scale[n] + col_shift[n]); an operand X is K-major
// (x[row * ld + k]) or MN-major (x[k * ld + row]).
struct Gemm3xDesc {
  const float* a; int64_t lda; int a_mn;
  const float* b; int64_t ldb; int b_mn;
  float* c; int64_t ldc;
  int m, n, k;
  float alpha; const float* col_scale; const float* col_shift; int relu;
};
bool gemm3x_supported(const Gemm3xDesc& d);
int gemm3x_plan_ksplit(const Gemm3xDesc* d, int count);                       // K split that fills the SMs (1 = none)
size_t gemm3x_partial_floats(const Gemm3xDesc* d, int count, int ksplit);      // floats of the [ksplit][m][n] partial tiles
size_t gemm3x_stream_floats(const Gemm3xDesc* d, int count);                   // floats of the stream mapping's shared-tile slots
int gemm3x_launch(cudaStream_t st, const Gemm3xDesc* d, int count, int ksplit, bool keep_partials, float* partial,
                  size_t partial_floats);

// finalize.cu: 3xTF32 operand preparation.  pattern 0 = [hi | lo | hi] (A side), 1 = [hi | hi | lo] (B side).
void launch_split_rows(const float* x, const int64_t* ids, int rows, int cols, int64_t ld_x, float scale, int pattern,
                       float* out, cudaStream_t st);
void launch_split_transpose(const float* x, int rows, int cols, int64_t ld_x, float scale, int pattern, float* out,
                            cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// Total order used everywhere: (score descending, post index ascending).
// A candidate is packed into one u64 key so that "larger key" == "ranks earlier":
//   high 32 bits: order-preserving transform of the fp32 score (-0.0 folded into +0.0 so that the
//                 two zeros tie, as they do under Python's float comparison at evaluator.py:109)
//   low  32 bits: ~index  (smaller index -> larger key)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t score_to_ordered(float s) {
  s = s + 0.0f;
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(s);
#else
  union { float f; uint32_t u; } cv; cv.f = s; uint32_t u = cv.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_score(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } cv; cv.u = u; return cv.f;
#endif
}
__host__ __device__ __forceinline__ unsigned long long make_key(float s, uint32_t index) {
  return ((unsigned long long)score_to_ordered(s) << 32) | (unsigned long long)(0xFFFFFFFFu - index);
}
__host__ __device__ __forceinline__ float key_score(unsigned long long k) { return ordered_to_score((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_index(unsigned long long k) { return 0xFFFFFFFFu - (uint32_t)k; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace frx
