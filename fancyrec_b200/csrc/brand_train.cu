// A4, training path: BrandAspects.forward + the mean over aspects (model.py:419-428, :594) WITH its dropout, forward
// and backward, without the [B, A, D] tensor the reference materialises (12.6 GB at B = 512, A = 2000, D = 3072):
//
//   out[b, d]  = (2 / A) * sum_a  w[b, a] * E[a, d] * m(b, a, d)            (dropout p = 0.5: keep -> x 2)
//   dW[b, a]   = (2 / A) * sum_d  g[b, d] * E[a, d] * m(b, a, d)  +  1e-4 * sign(w[b, a])     (L1Penalty.backward)
//   dE[a, d]   = (2 / A) * sum_b  g[b, d] * w[b, a] * m(b, a, d)
//
// The dropout acts on the B x A x D PRODUCTS, so none of the three is a GEMM: every term carries its own keep bit.  The
// bits are not stored either: m(b, a, d) is a counter-based hash of (seed, b, a, d), one 32-bit word per 32 elements,
// regenerated identically by the three kernels (and by frx_brand_dropout_mask, which writes it out for the tests).
// CUDA-core kernels, shared-memory bound (one conflict-free 4-byte shared load per masked FMA or per two).
//
// Element <-> bit: inside a 1024-wide column tile, d = tile * 1024 + i * 32 + j  (i, j in 0..31) is bit i of the word
// (b, a, tile, j): a thread that owns word j reads columns j, j + 32, ... -- consecutive lanes, consecutive words.
#include "common.cuh"

namespace frx {
namespace bt {

constexpr int DT = 1024;        // column tile
constexpr int ROWS = 16;        // rows (b, or a in the dE kernel) per block of the two "row x column" kernels
constexpr int KC = 8;           // reduction steps staged in shared memory at a time

__device__ __forceinline__ uint32_t mask_word(uint32_t seed_lo, uint32_t seed_hi, uint32_t b, uint32_t a, uint32_t n_a,
                                              uint32_t word) {
  uint32_t x = (b * n_a + a) * 0x9E3779B1u + word * 0x85EBCA77u + seed_lo;
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x += seed_hi; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}

// out[r, d] = scale * sum_k  L[r, k] * R[k, d] * m(., ., d)   for two flavours:
//   FWD: r = b, k = a, L = w [B, A] (ld_l),  R = E [A, D]                       -> out [B, D]
//   DE : r = a, k = b, L = w^T (element (a, b) = w[b * ld_l + a]), R = g [B, D]  -> dE  [A, D]
// Block = 256 threads = 8 warps; warp w owns rows 2w, 2w + 1 of the block's 16; lane j owns the 32 columns j + 32 i.
template <bool DE>
__global__ void __launch_bounds__(256) masked_rows_kernel(const float* __restrict__ lmat, int64_t ld_l,
                                                          const float* __restrict__ rmat, int n_rows, int n_k, int n_d,
                                                          int n_a, uint32_t seed_lo, uint32_t seed_hi, float scale,
                                                          float* __restrict__ out) {
  __shared__ float s_r[KC][DT];           // R[k0 .. k0 + KC, tile columns]
  __shared__ float s_l[ROWS][KC];         // L[rows, k0 .. k0 + KC]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, d0 = tile * DT;
  const int r0 = blockIdx.y * ROWS + warp * 2;
  float acc[2][32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { acc[0][i] = 0.f; acc[1][i] = 0.f; }
  for (int k0 = 0; k0 < n_k; k0 += KC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < KC * (DT / 4); idx += 256) {
      const int kk = idx / (DT / 4), c = (idx % (DT / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < n_k) {
        const float* src = rmat + (int64_t)(k0 + kk) * n_d + d0 + c;
        if (d0 + c + 3 < n_d && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (d0 + c < n_d) v.x = __ldg(src);
          if (d0 + c + 1 < n_d) v.y = __ldg(src + 1);
          if (d0 + c + 2 < n_d) v.z = __ldg(src + 2);
          if (d0 + c + 3 < n_d) v.w = __ldg(src + 3);
        }
      }
      *reinterpret_cast<float4*>(&s_r[kk][c]) = v;
    }
    if (threadIdx.x < ROWS * KC) {
      const int rr = threadIdx.x / KC, kk = threadIdx.x % KC;
      const int r = blockIdx.y * ROWS + rr, k = k0 + kk;
      float v = 0.f;
      if (r < n_rows && k < n_k) v = DE ? __ldg(lmat + (int64_t)k * ld_l + r) : __ldg(lmat + (int64_t)r * ld_l + k);
      s_l[rr][kk] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int kk = 0; kk < KC; ++kk) {
      const int k = k0 + kk;
      if (k >= n_k) break;
      // (b, a) of this step for each of the warp's two rows
      const uint32_t m0 = DE ? mask_word(seed_lo, seed_hi, (uint32_t)k, (uint32_t)r0, (uint32_t)n_a, (uint32_t)(tile * 32 + lane))
                             : mask_word(seed_lo, seed_hi, (uint32_t)r0, (uint32_t)k, (uint32_t)n_a, (uint32_t)(tile * 32 + lane));
      const uint32_t m1 = DE ? mask_word(seed_lo, seed_hi, (uint32_t)k, (uint32_t)(r0 + 1), (uint32_t)n_a, (uint32_t)(tile * 32 + lane))
                             : mask_word(seed_lo, seed_hi, (uint32_t)(r0 + 1), (uint32_t)k, (uint32_t)n_a, (uint32_t)(tile * 32 + lane));
      const float l0 = s_l[warp * 2][kk], l1 = s_l[warp * 2 + 1][kk];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = s_r[kk][i * 32 + lane];
        acc[0][i] = fmaf((m0 >> i) & 1u ? l0 : 0.f, e, acc[0][i]);
        acc[1][i] = fmaf((m1 >> i) & 1u ? l1 : 0.f, e, acc[1][i]);
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int r = r0 + rr;
    if (r >= n_rows) continue;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int d = d0 + i * 32 + lane;
      if (d < n_d) out[(int64_t)r * n_d + d] = scale * acc[rr][i];
    }
  }
}

// dW[b, a] = scale * sum_d g[b, d] * E[a, d] * m(b, a, d) + l1 * sign(w[b, a]).
// Block = 8 b x 8 a outputs, 256 threads: thread (bl, al, q) sums the words j = 8 q .. 8 q + 7 of every column tile.
constexpr int WB = 8, WA = 8;
__global__ void __launch_bounds__(256) masked_dw_kernel(const float* __restrict__ g, const float* __restrict__ e,
                                                        const float* __restrict__ w, int64_t ld_w, int n_b, int n_a, int n_d,
                                                        uint32_t seed_lo, uint32_t seed_hi, float scale, float l1,
                                                        float* __restrict__ dw) {
  constexpr int HALF = DT / 2;            // half a column tile at a time (bits 0..15, then 16..31, of every word)
  __shared__ float s_g[WB][HALF + 1];     // +1: rows start on different banks
  __shared__ float s_e[WA][HALF + 1];
  __shared__ float s_part[4][WB * WA];
  const int q = threadIdx.x >> 6, al = (threadIdx.x >> 3) & 7, bl = threadIdx.x & 7;
  const int b = blockIdx.y * WB + bl, a = blockIdx.x * WA + al;
  float acc = 0.f;
  const int n_tiles = (n_d + DT - 1) / DT;
  for (int th = 0; th < 2 * n_tiles; ++th) {
    const int tile = th >> 1, half = th & 1;
    const int d0 = tile * DT + half * HALF;
    __syncthreads();
    for (int idx = threadIdx.x; idx < (WB + WA) * HALF; idx += 256) {
      const int row = idx / HALF, c = idx % HALF;
      float v = 0.f;
      if (d0 + c < n_d) {
        if (row < WB) { const int bb = blockIdx.y * WB + row; if (bb < n_b) v = __ldg(g + (int64_t)bb * n_d + d0 + c); }
        else { const int aa = blockIdx.x * WA + row - WB; if (aa < n_a) v = __ldg(e + (int64_t)aa * n_d + d0 + c); }
      }
      if (row < WB) s_g[row][c] = v; else s_e[row - WB][c] = v;
    }
    __syncthreads();
    if (b < n_b && a < n_a) {
#pragma unroll 1
      for (int j = q * 8; j < q * 8 + 8; ++j) {
        const uint32_t m = mask_word(seed_lo, seed_hi, (uint32_t)b, (uint32_t)a, (uint32_t)n_a, (uint32_t)(tile * 32 + j)) >> (half * 16);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          acc = fmaf((m >> i) & 1u ? s_g[bl][i * 32 + j] : 0.f, s_e[al][i * 32 + j], acc);
      }
    }
  }
  s_part[q][al * WB + bl] = acc;
  __syncthreads();
  if (q == 0 && b < n_b && a < n_a) {
    const float total = s_part[0][al * WB + bl] + s_part[1][al * WB + bl] + s_part[2][al * WB + bl] + s_part[3][al * WB + bl];
    const float wv = __ldg(w + (int64_t)b * ld_w + a);
    const float sgn = wv > 0.f ? 1.f : (wv < 0.f ? -1.f : 0.f);       // torch.sign: 0 at 0 (and NaN stays out of scope)
    dw[(int64_t)b * n_a + a] = scale * total + l1 * sgn;
  }
}

__global__ void mask_dump_kernel(int n_b, int n_a, int n_d, uint32_t seed_lo, uint32_t seed_hi, uint8_t* __restrict__ out) {
  const int64_t total = (int64_t)n_b * n_a * n_d;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(idx % n_d);
    const int64_t ba = idx / n_d;
    const int a = (int)(ba % n_a), b = (int)(ba / n_a);
    const int tile = d / DT, within = d % DT, i = within / 32, j = within % 32;
    out[idx] = (uint8_t)((mask_word(seed_lo, seed_hi, (uint32_t)b, (uint32_t)a, (uint32_t)n_a, (uint32_t)(tile * 32 + j)) >> i) & 1u);
  }
}

}  // namespace bt
}  // namespace frx

extern "C" {

int frx_brand_train_fwd(const float* w_rows, int64_t ld_w, const float* e, int b, int a, int d, uint64_t seed, float* out,
                        void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(w_rows && e && out, "frx_brand_train_fwd: NULL pointer");
  FRX_CHECK_ARG(b > 0 && a > 0 && d > 0 && ld_w >= a, "frx_brand_train_fwd: bad sizes");
  int dev = 0;
  FRX_CUDA(cudaGetDevice(&dev));
  const int rc = frx_device_check(dev);
  if (rc) return rc;
  dim3 grid((d + bt::DT - 1) / bt::DT, (b + bt::ROWS - 1) / bt::ROWS);
  bt::masked_rows_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(w_rows, ld_w, e, b, a, d, a, (uint32_t)seed,
                                                                       (uint32_t)(seed >> 32), 2.0f / (float)a, out);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_brand_train_bwd(const float* grad_out, const float* w_rows, int64_t ld_w, const float* e, int b, int a, int d,
                        uint64_t seed, float* d_w_rows, float* d_e, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(grad_out && w_rows && e && d_w_rows && d_e, "frx_brand_train_bwd: NULL pointer");
  FRX_CHECK_ARG(b > 0 && a > 0 && d > 0 && ld_w >= a, "frx_brand_train_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = 2.0f / (float)a;
  const uint32_t lo = (uint32_t)seed, hi = (uint32_t)(seed >> 32);
  dim3 gw((a + bt::WA - 1) / bt::WA, (b + bt::WB - 1) / bt::WB);
  bt::masked_dw_kernel<<<gw, 256, 0, st>>>(grad_out, e, w_rows, ld_w, b, a, d, lo, hi, scale, 1e-4f, d_w_rows);
  FRX_LAUNCH_CHECK();
  dim3 ge((d + bt::DT - 1) / bt::DT, (a + bt::ROWS - 1) / bt::ROWS);
  bt::masked_rows_kernel<true><<<ge, 256, 0, st>>>(w_rows, ld_w, grad_out, a, b, d, a, lo, hi, scale, d_e);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_brand_dropout_mask(int b, int a, int d, uint64_t seed, uint8_t* mask, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(mask && b > 0 && a > 0 && d > 0, "frx_brand_dropout_mask: bad arguments");
  bt::mask_dump_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(b, a, d, (uint32_t)seed, (uint32_t)(seed >> 32), mask);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

}  // extern "C"
