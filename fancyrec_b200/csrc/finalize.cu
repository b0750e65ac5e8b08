// A1-A4: post-embedding finalisation (frame mean-pool -> per-branch l2norm -> concat -> row l2norm
// -> fp32 / bf16) as ONE pass over HBM, and the brand embedding (W.E)/A without the [NB,A,D]
// intermediate.  HBM-bound streaming kernels: 128-bit loads, one CTA per post row, the pooled row is
// staged in shared memory so every input byte is read exactly once and every output byte written once.
//
// Reference lines replaced: util/data_provider.py:40,91,132 (torch.mean(frames, 0)),
// model.py:39-44 (l2norm), model.py:482-485 (cat), evaluator.py:14-19,27-28, model.py:419-428,594.
#include "common.cuh"

namespace frx {

constexpr int kFinThreads = 256;

struct FinalizeParams {
  const float* visual;
  const int64_t* row_ptr;
  const int32_t* row_idx;
  const float* text;
  int64_t n_posts;
  int dv, dt, flags;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int64_t ld_bf16;
};

__device__ __forceinline__ float2 block_sum2(float a, float b, float* red /* [2*32] */) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // protect `red` from the previous use
  if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float x = lane < nw ? red[lane] : 0.f;
  float y = lane < nw ? red[32 + lane] : 0.f;
  x = warp_sum(x);
  y = warp_sum(y);
  return make_float2(x, y);
}

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  using T = float4;
  static __device__ __forceinline__ T ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ void add(T& a, const T& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
  static __device__ __forceinline__ void div(T& a, float d) { a.x /= d; a.y /= d; a.z /= d; a.w /= d; }
  static __device__ __forceinline__ float sq(const T& a) { return a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w; }
  static __device__ __forceinline__ void st(float* p, const T& a) { *reinterpret_cast<float4*>(p) = a; }
  static __device__ __forceinline__ T lds(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st_bf16(__nv_bfloat16* p, const T& a) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
template <>
struct Vec<1> {
  using T = float;
  static __device__ __forceinline__ T ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ T zero() { return 0.f; }
  static __device__ __forceinline__ void add(T& a, const T& b) { a += b; }
  static __device__ __forceinline__ void div(T& a, float d) { a /= d; }
  static __device__ __forceinline__ float sq(const T& a) { return a * a; }
  static __device__ __forceinline__ void st(float* p, const T& a) { *p = a; }
  static __device__ __forceinline__ T lds(const float* p) { return *p; }
  static __device__ __forceinline__ void st_bf16(__nv_bfloat16* p, const T& a) { *p = __float2bfloat16_rn(a); }
};

template <int VEC>
__global__ void __launch_bounds__(kFinThreads) finalize_kernel(FinalizeParams P) {
  extern __shared__ __align__(16) float row[];   // [dv + dt]
  __shared__ float red[64];
  using V = Vec<VEC>;
  const int d = P.dv + P.dt;
  const int tid = threadIdx.x;

  for (int64_t p = blockIdx.x; p < P.n_posts; p += gridDim.x) {
    // ---- pass A: global -> smem (mean-pool the visual branch), per-branch sum of squares ----
    float ssv = 0.f, sst = 0.f;
    int64_t r0 = p, r1 = p + 1;
    if (P.row_ptr) { r0 = P.row_ptr[p]; r1 = P.row_ptr[p + 1]; }
    const float nf = (float)(r1 - r0);
    for (int c = tid * VEC; c < P.dv; c += kFinThreads * VEC) {
      typename V::T acc = V::zero();
      int64_t r = r0;
      // 4 independent 128-bit loads in flight per thread; summed in frame order
      for (; r + 4 <= r1; r += 4) {
        int64_t i0 = r, i1 = r + 1, i2 = r + 2, i3 = r + 3;
        if (P.row_idx) { i0 = P.row_idx[i0]; i1 = P.row_idx[i1]; i2 = P.row_idx[i2]; i3 = P.row_idx[i3]; }
        typename V::T a0 = V::ld(P.visual + i0 * P.dv + c), a1 = V::ld(P.visual + i1 * P.dv + c);
        typename V::T a2 = V::ld(P.visual + i2 * P.dv + c), a3 = V::ld(P.visual + i3 * P.dv + c);
        V::add(acc, a0); V::add(acc, a1); V::add(acc, a2); V::add(acc, a3);
      }
      for (; r < r1; ++r) {
        int64_t i0 = P.row_idx ? (int64_t)P.row_idx[r] : r;
        typename V::T a0 = V::ld(P.visual + i0 * P.dv + c);
        V::add(acc, a0);
      }
      if (P.row_ptr) V::div(acc, nf);          // torch.mean on CPU = sum / F (0 frames -> NaN)
      ssv += V::sq(acc);
      V::st(row + c, acc);
    }
    for (int c = tid * VEC; c < P.dt; c += kFinThreads * VEC) {
      typename V::T a = V::ld(P.text + p * P.dt + c);
      sst += V::sq(a);
      V::st(row + P.dv + c, a);
    }
    float2 ss = block_sum2(ssv, sst, red);      // also orders the smem writes above
    // ---- pass B (only with per-branch norms): scale branches in smem, total sum of squares ----
    float total = ss.x + ss.y;
    const bool vn = (P.flags & FRX_VISUAL_NORM) != 0, tn = (P.flags & FRX_TEXT_NORM) != 0 && P.dt > 0;
    if (vn || tn) {
      const float nv = vn ? sqrtf(ss.x) : 1.f, nt = tn ? sqrtf(ss.y) : 1.f;
      float s2 = 0.f;
      for (int c = tid * VEC; c < d; c += kFinThreads * VEC) {
        typename V::T a = V::lds(row + c);
        const bool is_v = c < P.dv;
        if (is_v ? vn : tn) { V::div(a, is_v ? nv : nt); V::st(row + c, a); }   // own elements only
        s2 += V::sq(a);
      }
      total = block_sum2(s2, 0.f, red).x;
    }
    // ---- pass C: smem -> global ----
    const bool fn = (P.flags & FRX_FINAL_NORM) != 0;
    const float nrm = sqrtf(total);
    for (int c = tid * VEC; c < d; c += kFinThreads * VEC) {
      typename V::T a = V::lds(row + c);
      if (fn) V::div(a, nrm);
      if (P.out_f32) V::st(P.out_f32 + p * d + c, a);
      if (P.out_bf16) V::st_bf16(P.out_bf16 + p * P.ld_bf16 + c, a);
    }
    if (P.out_bf16) {
      for (int64_t c = d + tid; c < P.ld_bf16; c += kFinThreads) P.out_bf16[p * P.ld_bf16 + c] = __float2bfloat16_rn(0.f);
    }
    __syncthreads();   // row[] is reused by the next post
  }
}

// ---------------------------------------------------------------------------------------------
// Register-resident variant for rows of up to NV*128 floats (the configs' 1024 / 2048 / 3072):
// ONE WARP per post row, the whole row lives in registers (NV float4 per lane), every load of a row
// (or of one frame of it) is issued before the first use -> up to NV 128-bit loads in flight per lane,
// no shared memory, no block barriers; reductions are warp shuffles.  This is the HBM-roofline path.
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) finalize_warp_kernel(FinalizeParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int d = P.dv + P.dt;
  const bool vn = (P.flags & FRX_VISUAL_NORM) != 0, tn = (P.flags & FRX_TEXT_NORM) != 0 && P.dt > 0;
  const bool fn = (P.flags & FRX_FINAL_NORM) != 0;
  for (int64_t p = wid; p < P.n_posts; p += nwarps) {
    float4 x[NV];
    if (P.row_ptr == nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        x[i] = c < P.dv ? __ldg(reinterpret_cast<const float4*>(P.visual + p * P.dv + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const int64_t r0 = P.row_ptr[p], r1 = P.row_ptr[p + 1];
#pragma unroll
      for (int i = 0; i < NV; ++i) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t r = r0; r < r1; ++r) {           // frame order = the reference's summation order
        const int64_t src = P.row_idx ? (int64_t)P.row_idx[r] : r;
        float4 f[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = (i * 32 + lane) * 4;
          f[i] = c < P.dv ? __ldg(reinterpret_cast<const float4*>(P.visual + src * P.dv + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) { x[i].x += f[i].x; x[i].y += f[i].y; x[i].z += f[i].z; x[i].w += f[i].w; }
      }
      // mean = sum * (1/F): one IEEE division per row instead of one per element (<= 1 ulp from sum / F,
      // inside the stated 4e-6 tolerance; exact whenever F is a power of two)
      const float inv_f = 1.0f / (float)(r1 - r0);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if ((i * 32 + lane) * 4 < P.dv) { x[i].x *= inv_f; x[i].y *= inv_f; x[i].z *= inv_f; x[i].w *= inv_f; }
      }
    }
    if (P.dt > 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c >= P.dv && c < d) x[i] = __ldg(reinterpret_cast<const float4*>(P.text + p * P.dt + (c - P.dv)));
      }
    }
    float ssv = 0.f, sst = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float q = x[i].x * x[i].x + x[i].y * x[i].y + x[i].z * x[i].z + x[i].w * x[i].w;
      if (c < P.dv) ssv += q; else if (c < d) sst += q;
    }
    ssv = warp_sum(ssv);
    sst = warp_sum(sst);
    float total = ssv + sst;
    if (vn || tn) {
      const float iv = vn ? 1.0f / sqrtf(ssv) : 1.f, it = tn ? 1.0f / sqrtf(sst) : 1.f;
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) {
          const float sc = c < P.dv ? iv : it;
          x[i].x *= sc; x[i].y *= sc; x[i].z *= sc; x[i].w *= sc;
          s2 += x[i].x * x[i].x + x[i].y * x[i].y + x[i].z * x[i].z + x[i].w * x[i].w;
        }
      }
      total = warp_sum(s2);
    }
    const float inv_nrm = fn ? 1.0f / sqrtf(total) : 1.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        float4 a = x[i];
        a.x *= inv_nrm; a.y *= inv_nrm; a.z *= inv_nrm; a.w *= inv_nrm;
        if (P.out_f32) *reinterpret_cast<float4*>(P.out_f32 + p * d + c) = a;
        if (P.out_bf16) Vec<4>::st_bf16(P.out_bf16 + p * P.ld_bf16 + c, a);
      }
    }
    if (P.out_bf16) {
      for (int64_t c = d + lane * 4; c < P.ld_bf16; c += 128)   // zero padding (d and ld are multiples of 4)
        *reinterpret_cast<uint2*>(P.out_bf16 + p * P.ld_bf16 + c) = make_uint2(0u, 0u);
    }
  }
}

// Streaming variant for un-pooled rows (one visual row + optional text row per post): pass 1 reads the
// row in batches of 8 independent 128-bit loads and only accumulates the sums of squares; pass 2 reads
// it again (an L2 hit: the 12 KB row was fetched microseconds earlier), scales and writes.  DRAM traffic
// is unchanged, but the kernel needs ~40 registers instead of ~180, so 6x more warps are resident and
// the HBM pipe stays full while other warps are in their arithmetic / store phases.
__device__ __forceinline__ float4 ld_row_f4(const FinalizeParams& P, int64_t p, int c) {
  // plain read-only loads: evict-first streaming hints were measured slower here (3.40 vs 3.29 ms at config 2) and
  // break the L2 re-read of finalize_stream_kernel
  return c < P.dv ? __ldg(reinterpret_cast<const float4*>(P.visual + p * P.dv + c))
                  : __ldg(reinterpret_cast<const float4*>(P.text + p * P.dt + (c - P.dv)));
}

__global__ void __launch_bounds__(256) finalize_stream_kernel(FinalizeParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int d = P.dv + P.dt;
  const bool vn = (P.flags & FRX_VISUAL_NORM) != 0, tn = (P.flags & FRX_TEXT_NORM) != 0 && P.dt > 0;
  const bool fn = (P.flags & FRX_FINAL_NORM) != 0;
  constexpr int B = 8;
  for (int64_t p = wid; p < P.n_posts; p += nwarps) {
    float ssv = 0.f, sst = 0.f;
    for (int c0 = lane * 4; c0 < d; c0 += B * 128) {
      float4 f[B];
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const int c = c0 + j * 128;
        f[j] = c < d ? ld_row_f4(P, p, c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const float q = f[j].x * f[j].x + f[j].y * f[j].y + f[j].z * f[j].z + f[j].w * f[j].w;
        if (c0 + j * 128 < P.dv) ssv += q; else sst += q;
      }
    }
    ssv = warp_sum(ssv);
    sst = warp_sum(sst);
    // branch scales; the final norm of the scaled row is computed from the scaled sums
    const float iv = vn ? 1.0f / sqrtf(ssv) : 1.f, it = tn ? 1.0f / sqrtf(sst) : 1.f;
    float total = ssv + sst;
    // squared norm of the branch-scaled row: each branch contributes ss * scale^2 (== 1 up to rounding when
    // that branch is normalised); identical to re-summing the scaled elements to within 1e-7 relative
    if (vn || tn) total = ssv * iv * iv + sst * it * it;
    const float inv_nrm = fn ? 1.0f / sqrtf(total) : 1.f;
    for (int c0 = lane * 4; c0 < d; c0 += B * 128) {
      float4 f[B];
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const int c = c0 + j * 128;
        f[j] = c < d ? ld_row_f4(P, p, c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const int c = c0 + j * 128;
        if (c < d) {
          const float sc = (c < P.dv) ? iv : it;
          float4 a = f[j];
          a.x = a.x * sc * inv_nrm; a.y = a.y * sc * inv_nrm; a.z = a.z * sc * inv_nrm; a.w = a.w * sc * inv_nrm;
          if (P.out_f32) *reinterpret_cast<float4*>(P.out_f32 + p * d + c) = a;
          if (P.out_bf16) Vec<4>::st_bf16(P.out_bf16 + p * P.ld_bf16 + c, a);
        }
      }
    }
    if (P.out_bf16) {
      for (int64_t c = d + lane * 4; c < P.ld_bf16; c += 128)
        *reinterpret_cast<uint2*>(P.out_bf16 + p * P.ld_bf16 + c) = make_uint2(0u, 0u);
    }
  }
}

// Block-per-row variant for long un-pooled rows (1024 < D <= 4096; the config-2 shape): 256 threads hold
// the row in registers (VPT float4 each, 12-16 registers), ONE block reduction per row (double-buffered
// scratch -> a single __syncthreads), reciprocal scaling.  ~40 registers/thread -> 8 blocks = 64 warps per
// SM keep the HBM pipe full while other blocks are in their reduction / store phase.
template <int VPT, bool POOLED>
__global__ void __launch_bounds__(256) finalize_block_kernel(FinalizeParams P) {
  __shared__ float red[2][2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = P.dv + P.dt;
  const bool vn = (P.flags & FRX_VISUAL_NORM) != 0, tn = (P.flags & FRX_TEXT_NORM) != 0 && P.dt > 0;
  const bool fn = (P.flags & FRX_FINAL_NORM) != 0;
  int buf = 0;
  float4 nxt[VPT];                                    // un-pooled path: next row's loads are issued one row ahead
  if (!POOLED) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 256 + tid) * 4;
      nxt[i] = (blockIdx.x < P.n_posts && c < d) ? ld_row_f4(P, blockIdx.x, c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int64_t p = blockIdx.x; p < P.n_posts; p += gridDim.x, buf ^= 1) {
    float4 x[VPT];
    float ssv = 0.f, sst = 0.f;
    if (!POOLED) {
      const int64_t pn = p + gridDim.x;
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int c = (i * 256 + tid) * 4;
        x[i] = nxt[i];
        nxt[i] = (pn < P.n_posts && c < d) ? ld_row_f4(P, pn, c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      // pooled visual branch: frames summed in order, 4 frames x (visual float4 of this thread) loads in flight
      const int64_t r0 = P.row_ptr[p], r1 = P.row_ptr[p + 1];
#pragma unroll
      for (int i = 0; i < VPT; ++i) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      int64_t r = r0;
      for (; r + 4 <= r1; r += 4) {
        int64_t src[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) src[u] = P.row_idx ? (int64_t)P.row_idx[r + u] : r + u;
        float4 f[4][VPT];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < VPT; ++i) {
            const int c = (i * 256 + tid) * 4;
            f[u][i] = c < P.dv ? __ldg(reinterpret_cast<const float4*>(P.visual + src[u] * P.dv + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < VPT; ++i) { x[i].x += f[u][i].x; x[i].y += f[u][i].y; x[i].z += f[u][i].z; x[i].w += f[u][i].w; }
      }
      for (; r < r1; ++r) {
        const int64_t src = P.row_idx ? (int64_t)P.row_idx[r] : r;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          const int c = (i * 256 + tid) * 4;
          if (c < P.dv) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(P.visual + src * P.dv + c));
            x[i].x += f.x; x[i].y += f.y; x[i].z += f.z; x[i].w += f.w;
          }
        }
      }
      const float inv_f = 1.0f / (float)(r1 - r0);
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int c = (i * 256 + tid) * 4;
        if (c < P.dv) { x[i].x *= inv_f; x[i].y *= inv_f; x[i].z *= inv_f; x[i].w *= inv_f; }
        else if (c < d) x[i] = __ldg(reinterpret_cast<const float4*>(P.text + p * P.dt + (c - P.dv)));
      }
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 256 + tid) * 4;
      const float q = x[i].x * x[i].x + x[i].y * x[i].y + x[i].z * x[i].z + x[i].w * x[i].w;
      if (c < P.dv) ssv += q; else sst += q;
    }
    ssv = warp_sum(ssv);
    sst = warp_sum(sst);
    if (lane == 0) { red[buf][0][warp] = ssv; red[buf][1][warp] = sst; }
    __syncthreads();
    float tv = 0.f, tt = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tv += red[buf][0][w]; tt += red[buf][1][w]; }
    const float iv = vn ? 1.0f / sqrtf(tv) : 1.f, it = tn ? 1.0f / sqrtf(tt) : 1.f;
    const float total = (vn || tn) ? tv * iv * iv + tt * it * it : tv + tt;
    const float inv_nrm = fn ? 1.0f / sqrtf(total) : 1.f;
    const float sv = iv * inv_nrm, st = it * inv_nrm;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 256 + tid) * 4;
      if (c < d) {
        const float sc = c < P.dv ? sv : st;
        float4 a = x[i];
        a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc;
        if (P.out_f32) *reinterpret_cast<float4*>(P.out_f32 + p * d + c) = a;
        if (P.out_bf16) Vec<4>::st_bf16(P.out_bf16 + p * P.ld_bf16 + c, a);
      }
    }
    if (P.out_bf16) {
      for (int64_t c = d + tid * 4; c < P.ld_bf16; c += 1024)
        *reinterpret_cast<uint2*>(P.out_bf16 + p * P.ld_bf16 + c) = make_uint2(0u, 0u);
    }
  }
}

template <int NV>
static void launch_finalize_warp(const FinalizeParams& P, cudaStream_t st) {
  const int64_t warps_needed = P.n_posts;
  // whole blocks per SM, measured (tools/gpu_finalize_grid_warp.py): un-pooled rows of <= 1024 columns run best with
  // 2 blocks per SM (6.9 TB/s against 6.5 TB/s with 8); pooled rows are flat from 4 per SM upwards
  const int per_sm = (NV <= 8 && P.row_ptr == nullptr) ? 2 : 8;
  const int64_t max_blocks = (int64_t)num_sms() * per_sm;
  int64_t blocks = (warps_needed + 7) / 8;
  if (blocks > max_blocks) blocks = max_blocks;
  finalize_warp_kernel<NV><<<(int)blocks, 256, 0, st>>>(P);
}

// ---------------------------------------------------------------------------------------------
// Brand embedding: out[i, :] = (1/A) sum_a W[ids[i], a] * E[a, :]   fp32 FMA, 64x64x16 smem tiles.
// NB is at most ~10k and A = 2000, D <= 3072: at most 1.2e11 flop, a few ms once per evaluation.
// ---------------------------------------------------------------------------------------------
constexpr int kBM = 64, kBN = 64, kBK = 16;

__global__ void __launch_bounds__(256) brand_embed_kernel(const float* __restrict__ w, const float* __restrict__ e,
                                                           const int64_t* __restrict__ ids, int nb, int a, int d,
                                                           float* __restrict__ out) {
  __shared__ float sw[kBK][kBM + 4];   // W tile, transposed: [k][m]
  __shared__ float se[kBK][kBN + 4];   // E tile: [k][n]
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4x4 outputs each
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a; k0 += kBK) {
    // W tile: 64 rows x 16 k -> 1024 elements / 256 threads = 4 each
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = threadIdx.x + i * 256;
      int m = idx >> 4, k = idx & 15;
      float v = 0.f;
      if (m0 + m < nb && k0 + k < a) {
        int64_t r = ids ? ids[m0 + m] : (int64_t)(m0 + m);
        v = __ldg(w + r * a + k0 + k);
      }
      sw[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = threadIdx.x + i * 256;
      int k = idx >> 6, n = idx & 63;
      float v = 0.f;
      if (k0 + k < a && n0 + n < d) v = __ldg(e + (int64_t)(k0 + k) * d + n0 + n);
      se[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float wa[4], eb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) wa[i] = sw[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) eb[j] = se[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wa[i], eb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float fa = (float)a;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < nb && n < d) out[(int64_t)m * d + n] = acc[i][j] / fa;   // .mean(0) = sum / A
    }
}

// Faster variant when A % 8 == 0 and D % 4 == 0 (the usual 2000 x {1024, 2048, 3072}): 128x128x8 tiles,
// 8x8 outputs per thread, 128-bit shared-memory reads, register-staged double buffering.
constexpr int kFM = 128, kFN = 128, kFK = 8;

__global__ void __launch_bounds__(256) brand_embed_kernel_128(const float* __restrict__ w, const float* __restrict__ e,
                                                               const int64_t* __restrict__ ids, int nb, int a, int d,
                                                               float* __restrict__ out) {
  __shared__ __align__(16) float sw[2][kFK][kFM];   // W tile, k-major
  __shared__ __align__(16) float se[2][kFK][kFN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kFM, n0 = blockIdx.x * kFN;
  const int ty = tid >> 4, tx = tid & 15;
  // global -> register staging: W: row (tid >> 1), k offset (tid & 1) * 4 ; E: k (tid >> 5), col (tid & 31) * 4
  const int wm = tid >> 1, wk = (tid & 1) * 4;
  const int ek = tid >> 5, en = (tid & 31) * 4;
  const bool w_ok = m0 + wm < nb;
  const int64_t wrow = w_ok ? (ids ? ids[m0 + wm] : (int64_t)(m0 + wm)) : 0;
  const float* wp = w + wrow * a + wk;
  const bool e_ok = n0 + en < d;
  const float* ep = e + (int64_t)ek * d + n0 + en;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float4 rw = w_ok ? __ldg(reinterpret_cast<const float4*>(wp)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 re = e_ok ? __ldg(reinterpret_cast<const float4*>(ep)) : make_float4(0.f, 0.f, 0.f, 0.f);
  sw[0][wk + 0][wm] = rw.x; sw[0][wk + 1][wm] = rw.y; sw[0][wk + 2][wm] = rw.z; sw[0][wk + 3][wm] = rw.w;
  *reinterpret_cast<float4*>(&se[0][ek][en]) = re;
  __syncthreads();
  const int nk = a / kFK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      rw = w_ok ? __ldg(reinterpret_cast<const float4*>(wp + (kt + 1) * kFK)) : make_float4(0.f, 0.f, 0.f, 0.f);
      re = e_ok ? __ldg(reinterpret_cast<const float4*>(ep + (int64_t)(kt + 1) * kFK * d)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kFK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sw[cur][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sw[cur][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&se[cur][k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&se[cur][k][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      const int nxt = cur ^ 1;
      sw[nxt][wk + 0][wm] = rw.x; sw[nxt][wk + 1][wm] = rw.y; sw[nxt][wk + 2][wm] = rw.z; sw[nxt][wk + 3][wm] = rw.w;
      *reinterpret_cast<float4*>(&se[nxt][ek][en]) = re;
    }
    __syncthreads();
  }
  const float fa = (float)a;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= nb) continue;
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const int n = n0 + tx * 8 + j;
      if (n < d)
        *reinterpret_cast<float4*>(out + (int64_t)m * d + n) =
            make_float4(acc[i][j] / fa, acc[i][j + 1] / fa, acc[i][j + 2] / fa, acc[i][j + 3] / fa);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 3xTF32 brand embedding on the tensor cores: x = hi + lo with hi, lo exactly representable in tf32
// (10-bit mantissa), and  W.E = W_hi.E_hi + W_lo.E_hi + W_hi.E_lo  (+ O(2^-22)) is ONE tf32 GEMM over the
// K-concatenated operands  W' = [W_hi | W_lo | W_hi]  (nb x 3A)  and  E'^T = [E_hi | E_hi | E_lo]^T  (D x 3A),
// fp32 accumulation in TMEM: fp32-grade result (~1e-6 relative) at tensor-core speed.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0x00000FFFu + ((u >> 13) & 1u);          // round to nearest even at bit 13
  return __uint_as_float(u & 0xFFFFE000u);
}

// pattern 0 ("A side"): [hi | lo | hi]     pattern 1 ("B side"): [hi | hi | lo]
// A'.B'^T over the 3x wide K then equals a_hi.b_hi + a_lo.b_hi + a_hi.b_lo.
// rows of X (optionally gathered through ids), optionally pre-scaled -> split row of width 3 * cols
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ ids, int cols,
                                                         int64_t ld_x, float scale, int pattern, float* __restrict__ out) {
  const int r = blockIdx.x;
  const int64_t src = ids ? ids[r] : (int64_t)r;
  float* o = out + (int64_t)r * 3 * cols;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = x[src * ld_x + c] * scale;
    const float hi = tf32_round(v), lo = tf32_round(v - hi);
    o[c] = hi;
    o[cols + c] = pattern ? hi : lo;
    o[2 * cols + c] = pattern ? lo : hi;
  }
}

// X [rows, cols] -> transposed and split: out[n, :] (width 3 * rows) from column n of X   (32x32 smem transpose)
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ x, int rows, int cols, int64_t ld_x,
                                                              float scale, int pattern, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int k = k0 + ty + i, n = n0 + tx;
    tile[ty + i][tx] = (k < rows && n < cols) ? x[(int64_t)k * ld_x + n] * scale : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int n = n0 + ty + i, k = k0 + tx;
    if (n < cols && k < rows) {
      const float v = tile[tx][ty + i];
      const float hi = tf32_round(v), lo = tf32_round(v - hi);
      float* o = out + (int64_t)n * 3 * rows;
      o[k] = hi;
      o[rows + k] = pattern ? hi : lo;
      o[2 * rows + k] = pattern ? lo : hi;
    }
  }
}

// Same split for LONG operands (the post side of the score contraction, millions of rows): 128-bit loads and
// stores, grid-stride over rows, one warp per row segment.  cols % 4 == 0, 16-byte aligned pointers.
__global__ void __launch_bounds__(256) split_rows_vec_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ld_x,
                                                             int pattern, float* __restrict__ out) {
  const int c4 = cols >> 2;
  const int64_t total = rows * (int64_t)c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int c = (int)(i - r * c4) << 2;
    const float4 v = __ldcs(reinterpret_cast<const float4*>(x + r * ld_x + c));
    float4 hi, lo;
    hi.x = tf32_round(v.x); lo.x = tf32_round(v.x - hi.x);
    hi.y = tf32_round(v.y); lo.y = tf32_round(v.y - hi.y);
    hi.z = tf32_round(v.z); lo.z = tf32_round(v.z - hi.z);
    hi.w = tf32_round(v.w); lo.w = tf32_round(v.w - hi.w);
    float* o = out + r * 3 * (int64_t)cols + c;
    *reinterpret_cast<float4*>(o) = hi;
    *reinterpret_cast<float4*>(o + cols) = pattern ? hi : lo;
    *reinterpret_cast<float4*>(o + 2 * cols) = pattern ? lo : hi;
  }
}

// ---------------------------------------------------------------------------------------------
// Masked soft-max weighted pool: the per-sample loop of MultiHeadSelfAttention.forward
// (model.py:105-114)  weight[i, :len_i] = softmax(atten[i, :len_i]);  out = (weight * x).mean(dim=1)
// as one pass: block b soft-maxes its <= T logits in shared memory, then every thread accumulates its columns over
// the valid steps only (padded steps are never read).  The mean runs over the PADDED length T, as the reference's does.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_pool_kernel(const float* __restrict__ x, const float* __restrict__ logit,
                                                            const int64_t* __restrict__ lengths, int t_max, int d,
                                                            float* __restrict__ out) {
  extern __shared__ float wgt[];                 // [t_max]
  __shared__ float red[8];
  const int b = blockIdx.x;
  int64_t len64 = lengths[b];
  const int len = len64 < 0 ? 0 : (len64 > t_max ? t_max : (int)len64);
  const float* lg = logit + (int64_t)b * t_max;
  float m = -INFINITY;
  for (int t = threadIdx.x; t < len; t += blockDim.x) m = fmaxf(m, lg[t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  for (int w = 0; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float z = 0.f;
  for (int t = threadIdx.x; t < len; t += blockDim.x) { const float e = expf(lg[t] - m); wgt[t] = e; z += e; }
  z = warp_sum(z);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = z;
  __syncthreads();
  z = 0.f;
  for (int w = 0; w < 8; ++w) z += red[w];
  const float inv = len > 0 ? 1.0f / (z * (float)t_max) : 0.f;     // an empty sequence gives a zero row (all weights 0)
  const float* xb = x + (int64_t)b * t_max * d;
  float* ob = out + (int64_t)b * d;
  if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int t = 0;
      for (; t + 4 <= len; t += 4) {               // four independent 128-bit loads in flight
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(xb + (int64_t)(t + u) * d + c));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float w = wgt[t + u];
          acc.x = fmaf(w, v[u].x, acc.x); acc.y = fmaf(w, v[u].y, acc.y);
          acc.z = fmaf(w, v[u].z, acc.z); acc.w = fmaf(w, v[u].w, acc.w);
        }
      }
      for (; t < len; ++t) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(xb + (int64_t)t * d + c));
        const float w = wgt[t];
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      *reinterpret_cast<float4*>(ob + c) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
  } else {
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      float acc = 0.f;
      for (int t = 0; t < len; ++t) acc = fmaf(wgt[t], xb[(int64_t)t * d + c], acc);
      ob[c] = acc * inv;
    }
  }
}

void launch_split_rows(const float* x, const int64_t* ids, int rows, int cols, int64_t ld_x, float scale, int pattern,
                       float* out, cudaStream_t st) {
  split_rows_kernel<<<rows, 256, 0, st>>>(x, ids, cols, ld_x, scale, pattern, out);
}
void launch_split_transpose(const float* x, int rows, int cols, int64_t ld_x, float scale, int pattern, float* out,
                            cudaStream_t st) {
  dim3 tg((cols + 31) / 32, (rows + 31) / 32);
  split_transpose_kernel<<<tg, 256, 0, st>>>(x, rows, cols, ld_x, scale, pattern, out);
}

static size_t brand_ws_bytes(int nb, int a, int d) {
  return (((size_t)nb * 3 * a * sizeof(float) + 255) & ~(size_t)255) + (((size_t)d * 3 * a * sizeof(float) + 255) & ~(size_t)255);
}

}  // namespace frx

extern "C" {

int frx_finalize_posts(const float* visual, const int64_t* row_ptr, const int32_t* row_idx, const float* text,
                       int64_t n_posts, int dv, int dt, int flags, float* out_f32, uint16_t* out_bf16,
                       int64_t ld_bf16, void* stream) {
  return frx_finalize_posts_bounded(visual, row_ptr, row_idx, text, n_posts, dv, dt, flags, out_f32, out_bf16, ld_bf16, 0,
                                    stream);
}

int frx_finalize_posts_bounded(const float* visual, const int64_t* row_ptr, const int32_t* row_idx, const float* text,
                               int64_t n_posts, int dv, int dt, int flags, float* out_f32, uint16_t* out_bf16,
                               int64_t ld_bf16, int blocks_per_sm, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(blocks_per_sm >= 0 && blocks_per_sm <= 32, "frx_finalize_posts: blocks_per_sm %d outside 0..32", blocks_per_sm);
  if (n_posts == 0) return FRX_OK;
  FRX_CHECK_ARG(visual != nullptr && dv > 0, "frx_finalize_posts: visual is NULL or dv <= 0");
  FRX_CHECK_ARG(n_posts >= 0 && dt >= 0, "frx_finalize_posts: negative size");
  FRX_CHECK_ARG((dt == 0) == (text == nullptr), "frx_finalize_posts: text pointer and dt disagree");
  FRX_CHECK_ARG(row_idx == nullptr || row_ptr != nullptr, "frx_finalize_posts: row_idx needs row_ptr");
  FRX_CHECK_ARG(out_f32 || out_bf16, "frx_finalize_posts: no output requested");
  const int d = dv + dt;
  if (out_bf16) FRX_CHECK_ARG(ld_bf16 >= d && ld_bf16 % 8 == 0, "frx_finalize_posts: ld_bf16 must be >= dv+dt and a multiple of 8");
  if (n_posts == 0) return FRX_OK;
  const size_t smem = (size_t)d * sizeof(float);
  FRX_CHECK_ARG(smem <= 200 * 1024, "frx_finalize_posts: dv+dt = %d exceeds the 51200-column row cache", d);
  FinalizeParams P{visual, row_ptr, row_idx, text, n_posts, dv, dt, flags, out_f32,
                   reinterpret_cast<__nv_bfloat16*>(out_bf16), ld_bf16};
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = dv % 4 == 0 && dt % 4 == 0 && aligned16(visual) && aligned16(text) && aligned16(out_f32) &&
                   aligned16(out_bf16);
  int grid = (int)(n_posts < (int64_t)num_sms() * 8 ? n_posts : (int64_t)num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && (d > 2048 || (d > 1024 && row_ptr == nullptr)) && d <= 4096 && ld_bf16 % 4 == 0) {   // long rows: block per row
    // One resident wave of whole blocks per SM, measured per variant (tools/gpu_finalize_grid.py: bandwidth vs blocks per
    // SM): 4 per SM for the 3072-wide un-pooled rows (2.73 ms against 2.96 ms with 8 per SM queued as two partial waves),
    // 5 per SM for the 2048-wide rows and the pooled rows.  Grid-stride loops cover the rest.
    const bool wide_unpooled = row_ptr == nullptr && d > 2048;
    int64_t blocks = n_posts;
    const int64_t max_blocks = (int64_t)num_sms() * (blocks_per_sm > 0 ? blocks_per_sm : (wide_unpooled || d > 3072 ? 4 : 5));
    if (blocks > max_blocks) blocks = max_blocks;
    {
      // A bounded launch is meant to run NEXT TO a resident contraction CTA, and a block can only join an SM whose
      // shared-memory / L1 split matches its own: the contraction holds the maximum shared-memory carve-out, so ask for the
      // same split (measured: without it the two kernels never co-reside -- the SM has to drain to change the split).  The
      // default launch goes back to the default split (streaming loads are ~20 % faster with a normal L1).
      static int carve_state = 0;                                 // 0 = default split, 1 = max shared
      const int want = blocks_per_sm > 0 ? 1 : 0;
      if (want != carve_state) {
        const int v = want ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault;
        cudaFuncSetAttribute(finalize_block_kernel<2, false>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
        cudaFuncSetAttribute(finalize_block_kernel<3, false>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
        cudaFuncSetAttribute(finalize_block_kernel<4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
        cudaFuncSetAttribute(finalize_block_kernel<3, true>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
        cudaFuncSetAttribute(finalize_block_kernel<4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
        cudaGetLastError();
        carve_state = want;
      }
    }
    if (row_ptr != nullptr) {
      if (d <= 3072) finalize_block_kernel<3, true><<<(int)blocks, 256, 0, st>>>(P);
      else finalize_block_kernel<4, true><<<(int)blocks, 256, 0, st>>>(P);
    } else {
      if (d <= 2048) finalize_block_kernel<2, false><<<(int)blocks, 256, 0, st>>>(P);
      else if (d <= 3072) finalize_block_kernel<3, false><<<(int)blocks, 256, 0, st>>>(P);
      else finalize_block_kernel<4, false><<<(int)blocks, 256, 0, st>>>(P);
    }
  } else if (vec && row_ptr == nullptr && d > 1024 && ld_bf16 % 4 == 0) {
    int64_t blocks = (n_posts + 7) / 8;
    const int64_t max_blocks = (int64_t)num_sms() * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    finalize_stream_kernel<<<(int)blocks, 256, 0, st>>>(P);
  } else if (vec && d <= 4096 && ld_bf16 % 4 == 0) {
    if (d <= 1024) launch_finalize_warp<8>(P, st);
    else if (d <= 2048) launch_finalize_warp<16>(P, st);
    else if (d <= 3072) launch_finalize_warp<24>(P, st);
    else launch_finalize_warp<32>(P, st);
  } else if (vec) {
    if (smem > 48 * 1024) FRX_CUDA(cudaFuncSetAttribute(finalize_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    finalize_kernel<4><<<grid, kFinThreads, smem, st>>>(P);
  } else {
    if (smem > 48 * 1024) FRX_CUDA(cudaFuncSetAttribute(finalize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    finalize_kernel<1><<<grid, kFinThreads, smem, st>>>(P);
  }
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_softmax_pool(const float* x, const float* logits, const int64_t* lengths, int batch, int t_max, int d, float* out,
                     void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(x && logits && lengths && out, "frx_softmax_pool: NULL pointer");
  FRX_CHECK_ARG(batch >= 0 && t_max > 0 && d > 0, "frx_softmax_pool: bad sizes batch=%d t_max=%d d=%d", batch, t_max, d);
  FRX_CHECK_ARG((size_t)t_max * sizeof(float) <= 160 * 1024, "frx_softmax_pool: t_max = %d exceeds 40960 steps", t_max);
  int dev = 0;
  FRX_CUDA(cudaGetDevice(&dev));
  int rc = frx_device_check(dev);
  if (rc) return rc;
  if (batch == 0) return FRX_OK;
  const size_t smem = (size_t)t_max * sizeof(float);
  if (smem > 48 * 1024) FRX_CUDA(cudaFuncSetAttribute(softmax_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  softmax_pool_kernel<<<batch, 256, smem, (cudaStream_t)stream>>>(x, logits, lengths, t_max, d, out);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_split_tf32x3(const float* x, int64_t rows, int cols, int64_t ld_x, int side, float* out, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(x && out, "frx_split_tf32x3: NULL pointer");
  FRX_CHECK_ARG(rows >= 0 && cols > 0 && ld_x >= cols, "frx_split_tf32x3: bad sizes rows=%lld cols=%d", (long long)rows, cols);
  FRX_CHECK_ARG(side == 0 || side == 1, "frx_split_tf32x3: side must be 0 (brand) or 1 (post)");
  FRX_CHECK_ARG(cols % 4 == 0 && ld_x % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0,
                "frx_split_tf32x3: cols and ld_x must be multiples of 4 and the pointers 16-byte aligned");
  int dev = 0;
  FRX_CUDA(cudaGetDevice(&dev));
  int rc = frx_device_check(dev);
  if (rc) return rc;
  if (rows == 0) return FRX_OK;
  const int64_t total = rows * (int64_t)(cols / 4);
  int64_t blocks = (total + 255) / 256;
  const int64_t max_blocks = (int64_t)num_sms() * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  split_rows_vec_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld_x, side, out);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

size_t frx_brand_embed_workspace_bytes(int nb, int a, int d) {
  if (nb <= 0 || a <= 0 || d <= 0 || a % 4 != 0) return 0;     // 0: only the CUDA-core path is available
  return frx::brand_ws_bytes(nb, a, d);
}

int frx_brand_embed(const float* w, int64_t w_rows, const float* e, const int64_t* brand_ids, int nb, int a, int d,
                    float* out_f32, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(w && e && out_f32, "frx_brand_embed: NULL pointer");
  FRX_CHECK_ARG(nb >= 0 && a > 0 && d > 0, "frx_brand_embed: bad sizes nb=%d a=%d d=%d", nb, a, d);
  FRX_CHECK_ARG(brand_ids != nullptr || nb <= w_rows, "frx_brand_embed: nb=%d exceeds table rows %lld", nb, (long long)w_rows);
  if (nb == 0) return FRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // tensor-core path (3xTF32) when the caller provides scratch and the problem is worth a GEMM launch
  if (workspace != nullptr && a % 4 == 0 && workspace_bytes >= brand_ws_bytes(nb, a, d) &&
      (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (long)nb * d >= 128 * 256) {
    float* wsplit = reinterpret_cast<float*>(workspace);
    float* esplit = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) +
                                             (((size_t)nb * 3 * a * sizeof(float) + 255) & ~(size_t)255));
    launch_split_rows(w, brand_ids, nb, a, a, 1.0f, 0, wsplit, st);              // W'   = [W_hi | W_lo | W_hi]
    launch_split_transpose(e, a, d, d, 1.0f, 1, esplit, st);                     // E'^T = [E_hi | E_hi | E_lo]^T
    FRX_LAUNCH_CHECK();
    return dense_tf32_scaled(wsplit, 3 * (int64_t)a, esplit, 3 * (int64_t)a, nb, d, 3 * a, out_f32, d, 1.0f / (float)a, stream);
  }
  const bool aligned = (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(e) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0;
  if (a % kFK == 0 && d % 4 == 0 && aligned) {
    dim3 grid((d + kFN - 1) / kFN, (nb + kFM - 1) / kFM);
    brand_embed_kernel_128<<<grid, 256, 0, st>>>(w, e, brand_ids, nb, a, d, out_f32);
  } else {
    dim3 grid((d + kBN - 1) / kBN, (nb + kBM - 1) / kBM);
    brand_embed_kernel<<<grid, 256, 0, st>>>(w, e, brand_ids, nb, a, d, out_f32);
  }
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

}  // extern "C"
