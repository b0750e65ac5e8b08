// A12 / A13: the in-batch similarity-tile losses, forward + backward (loss.py:87-143,
// loss_ctrs.py:179-214).  fp32 end to end (the reference computes the tile in fp32); the B x B tile,
// its rank weights, the hinge / soft-max terms and dS are produced on the device without the B GEMV
// launches (loss.py:91-93), the four sorts (loss.py:96-105) or the B^2 host loop (loss.py:116-119) of
// the reference.  The GEMMs (tile, dS.brand, dS^T.post, softmax logits) run on the tensor cores as 3xTF32
// (fp32-grade) through the same tcgen05 kernel as the scoring path; small fused row kernels do the rest.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace frx {

// ---------------------------------------------------------------------------------------------
// C[M,N] = alpha * sum_k A(m,k) B(k,n) + beta * C,  A(m,k) = a[m*sam + k*sak], B(k,n) = b[k*sbk + n*sbn]
// 64x64x16 tiles, 256 threads, 4x4 outputs per thread.
// ---------------------------------------------------------------------------------------------
constexpr int GM = 64, GN = 64, GK = 16;

__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ a, int64_t sam, int64_t sak,
                                                     const float* __restrict__ b, int64_t sbk, int64_t sbn,
                                                     float* __restrict__ c, int64_t ldc, int m, int n, int k,
                                                     float alpha, float beta) {
  __shared__ float sa[GK][GM + 4];
  __shared__ float sb[GK][GN + 4];
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool a_kfast = sak == 1, b_kfast = sbk == 1;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < k; k0 += GK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      int mm, kk;
      if (a_kfast) { mm = idx >> 4; kk = idx & 15; } else { kk = idx >> 6; mm = idx & 63; }
      float v = 0.f;
      if (m0 + mm < m && k0 + kk < k) v = __ldg(a + (int64_t)(m0 + mm) * sam + (int64_t)(k0 + kk) * sak);
      sa[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      int nn, kk;
      if (b_kfast) { nn = idx >> 4; kk = idx & 15; } else { kk = idx >> 6; nn = idx & 63; }
      float v = 0.f;
      if (n0 + nn < n && k0 + kk < k) v = __ldg(b + (int64_t)(k0 + kk) * sbk + (int64_t)(n0 + nn) * sbn);
      sb[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sa[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sb[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int mm = m0 + ty * 4 + i, nn = n0 + tx * 4 + j;
      if (mm < m && nn < n) {
        float* dst = c + (int64_t)mm * ldc + nn;
        *dst = beta == 0.f ? alpha * acc[i][j] : alpha * acc[i][j] + beta * *dst;
      }
    }
}

static void sgemm(cudaStream_t st, const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbk, int64_t sbn,
                  float* c, int64_t ldc, int m, int n, int k, float alpha, float beta) {
  dim3 grid((n + GN - 1) / GN, (m + GM - 1) / GM);
  sgemm_kernel<<<grid, 256, 0, st>>>(a, sam, sak, b, sbk, sbn, c, ldc, m, n, k, alpha, beta);
}

// ---------------------------------------------------------------------------------------------
// rank weights of the diagonal (loss.py:96-105): block i counts, in row i and in column i, the entries
// that precede S[i,i] in a descending sort (ties: smaller index first).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_sum_int(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}
__device__ __forceinline__ float block_sum_float(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(128) tile_rank_kernel(const float* __restrict__ s, int b, float* __restrict__ rank_p,
                                                         float* __restrict__ rank_b, float* __restrict__ diag) {
  __shared__ int red[4];
  const int i = blockIdx.x;
  const float d = s[(int64_t)i * b + i];
  int cr = 0, cc = 0;
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    const float r = s[(int64_t)i * b + j], c = s[(int64_t)j * b + i];
    cr += (r > d) || (r == d && j < i);
    cc += (c > d) || (c == d && j < i);
  }
  cr = block_sum_int(cr, red);
  cc = block_sum_int(cc, red);
  if (threadIdx.x == 0) {
    const float fb = (float)b;
    rank_p[i] = 1.0f / (fb - (float)(cr + 1) + 1.0f) + 1.0f;
    rank_b[i] = 1.0f / (fb - (float)(cc + 1) + 1.0f) + 1.0f;
    diag[i] = d;
  }
}

// hinge + same-brand mask + column-broadcast weights (loss.py:107-132), loss partials and dS.
__global__ void __launch_bounds__(256) triplet_row_kernel(const float* __restrict__ s, const int64_t* __restrict__ ids, int b,
                                                           float margin, float scale, const float* __restrict__ rank_p,
                                                           const float* __restrict__ rank_b, const float* __restrict__ diag,
                                                           float* __restrict__ ds, float* __restrict__ partial) {
  __shared__ float redf[8];
  __shared__ int redi[8];
  const int i = blockIdx.x;
  const float di = diag[i];
  const int64_t idi = ids[i];
  float loss = 0.f, row_gp = 0.f;
  int col_cnt = 0;
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    const bool same = ids[j] == idi;
    const float sij = s[(int64_t)i * b + j];
    const float xp = margin + sij - di;         // cost_p argument (d1 = S[i,i])
    const float xb = margin + sij - diag[j];    // cost_b argument (d2 = S[j,j])
    float g = 0.f;
    if (!same) {
      const float wp = rank_p[j], wb = rank_b[j];
      loss += fmaxf(xp, 0.f) * wp + fmaxf(xb, 0.f) * wb;
      if (xp >= 0.f) { g += wp; row_gp += wp; }   // torch.clamp backward passes the gradient at x == min
      if (xb >= 0.f) g += wb;
    }
    if (j != i) ds[(int64_t)i * b + j] = scale * g;
    // column i: entries (j, i) whose cost_b argument margin + S[j,i] - S[i,i] is active
    const float xcol = margin + s[(int64_t)j * b + i] - di;
    col_cnt += (!same && xcol >= 0.f) ? 1 : 0;
  }
  loss = block_sum_float(loss, redf);
  row_gp = block_sum_float(row_gp, redf);
  col_cnt = block_sum_int(col_cnt, redi);
  if (threadIdx.x == 0) {
    partial[i] = loss;
    ds[(int64_t)i * b + i] = -scale * (row_gp + rank_b[i] * (float)col_cnt);
  }
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int n, float scale,
                                                               float* __restrict__ out) {
  __shared__ double red[8];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[0] = (float)(t * (double)scale);
  }
}

// ---------------------------------------------------------------------------------------------
// TripletLoss, everything between the tile GEMM and the gradient GEMMs in ONE cooperative launch (block i = row i):
//   phase 1  S[i,:] = sum over the K-split partial tiles of the tile GEMM, in split order (kept in shared memory and
//            written once to global memory for the column reads of the other blocks)
//   phase 2  rank of the diagonal in row i and in column i by counting (loss.py:96-105)  -> rank_p, rank_b, diag
//   phase 3  hinge, same-brand mask, column-broadcast weights (loss.py:107-132) -> dS[i,:], the row's loss
//   phase 4  block 0 adds the row losses in a fixed order (float64)
// Replaces reduce_ksplit + tile_rank + triplet_row + reduce_partials (4 launches, S read back from HBM three times).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) triplet_tile_kernel(const float* __restrict__ partial, int ksplit, int b,
                                                           const int64_t* __restrict__ ids, float margin, float scale,
                                                           float* __restrict__ s, float* __restrict__ rank_p,
                                                           float* __restrict__ rank_b, float* __restrict__ diag,
                                                           float* __restrict__ ds, float* __restrict__ row_loss,
                                                           float* __restrict__ loss_out) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float srow[];                 // S[i, 0..b)
  __shared__ float redf[8];
  __shared__ int redi[8];
  __shared__ double redd[8];
  const int i = blockIdx.x;
  const int64_t bb = (int64_t)b * b;
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    float acc = 0.f;
    for (int ks = 0; ks < ksplit; ++ks) acc += __ldcg(partial + (int64_t)ks * bb + (int64_t)i * b + j);
    srow[j] = acc;
    s[(int64_t)i * b + j] = acc;
  }
  grid.sync();
  const float di = srow[i];
  {
    int cr = 0, cc = 0;
    for (int j = threadIdx.x; j < b; j += blockDim.x) {
      const float r = srow[j], c = __ldcg(s + (int64_t)j * b + i);
      cr += (r > di) || (r == di && j < i);
      cc += (c > di) || (c == di && j < i);
    }
    cr = block_sum_int(cr, redi);
    cc = block_sum_int(cc, redi);
    if (threadIdx.x == 0) {
      const float fb = (float)b;
      rank_p[i] = 1.0f / (fb - (float)(cr + 1) + 1.0f) + 1.0f;
      rank_b[i] = 1.0f / (fb - (float)(cc + 1) + 1.0f) + 1.0f;
      diag[i] = di;
    }
  }
  grid.sync();
  {
    const int64_t idi = ids[i];
    float loss = 0.f, row_gp = 0.f;
    int col_cnt = 0;
    for (int j = threadIdx.x; j < b; j += blockDim.x) {
      const bool same = ids[j] == idi;
      const float sij = srow[j];
      const float xp = margin + sij - di;                       // cost_p argument (d1 = S[i,i])
      const float xb = margin + sij - __ldcg(diag + j);         // cost_b argument (d2 = S[j,j])
      float g = 0.f;
      if (!same) {
        const float wp = __ldcg(rank_p + j), wb = __ldcg(rank_b + j);
        loss += fmaxf(xp, 0.f) * wp + fmaxf(xb, 0.f) * wb;
        if (xp >= 0.f) { g += wp; row_gp += wp; }               // torch.clamp backward passes the gradient at x == min
        if (xb >= 0.f) g += wb;
      }
      if (ds != nullptr && j != i) ds[(int64_t)i * b + j] = scale * g;
      const float xcol = margin + __ldcg(s + (int64_t)j * b + i) - di;      // column i: cost_b arguments that use S[i,i]
      col_cnt += (!same && xcol >= 0.f) ? 1 : 0;
    }
    loss = block_sum_float(loss, redf);
    row_gp = block_sum_float(row_gp, redf);
    col_cnt = block_sum_int(col_cnt, redi);
    if (threadIdx.x == 0) {
      row_loss[i] = loss;
      if (ds != nullptr) ds[(int64_t)i * b + i] = -scale * (row_gp + __ldcg(rank_b + i) * (float)col_cnt);
    }
  }
  grid.sync();
  if (blockIdx.x == 0) {
    double v = 0.0;
    for (int r = threadIdx.x; r < b; r += blockDim.x) v += (double)__ldcg(row_loss + r);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) redd[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += redd[w];
      loss_out[0] = (float)(t * (double)scale);
    }
  }
}

// Can `blocks` blocks of the cooperative tile kernel be resident at once on this device?
static bool triplet_tile_fits(int blocks, size_t smem) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, triplet_tile_kernel, 256, smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return (long)per_sm * num_sms() >= blocks;
}

// ---------------------------------------------------------------------------------------------
// contrastive pieces
// ---------------------------------------------------------------------------------------------
// F.normalize: y = x / max(||x||, 1e-12); norm_out optional.
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ x, int d, float* __restrict__ y,
                                                              float* __restrict__ norm_out) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  float ss = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) { const float v = x[(int64_t)r * d + c]; ss += v * v; }
  ss = block_sum_float(ss, red);
  const float nrm = fmaxf(sqrtf(ss), 1e-12f);
  for (int c = threadIdx.x; c < d; c += blockDim.x) y[(int64_t)r * d + c] = x[(int64_t)r * d + c] / nrm;
  if (norm_out && threadIdx.x == 0) norm_out[r] = nrm;
}

// dx = (dy - y * <y, dy>) / norm   (backward of F.normalize away from the eps clamp)
__global__ void __launch_bounds__(256) normalize_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                             const float* __restrict__ nrm, int d, float* __restrict__ dx) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  float dot = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) dot += dy[(int64_t)r * d + c] * y[(int64_t)r * d + c];
  dot = block_sum_float(dot, red);
  const float n = nrm[r];
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    dx[(int64_t)r * d + c] = (dy[(int64_t)r * d + c] - y[(int64_t)r * d + c] * dot) / n;
}

// Both F.normalize calls of a loss in ONE launch (blocks [0, rows) -> x0, [rows, 2 rows) -> x1), 16-byte accesses when
// the rows allow.
__global__ void __launch_bounds__(256) normalize_rows2_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int rows,
                                                               int d, float* __restrict__ y0, float* __restrict__ y1,
                                                               float* __restrict__ n0, float* __restrict__ n1) {
  __shared__ float red[8];
  const bool second = (int)blockIdx.x >= rows;
  const int r = second ? (int)blockIdx.x - rows : (int)blockIdx.x;
  const float* x = (second ? x1 : x0) + (int64_t)r * d;
  float* y = (second ? y1 : y0) + (int64_t)r * d;
  const bool vec = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  float ss = 0.f;
  if (vec) {
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(x + c);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (int c = threadIdx.x; c < d; c += blockDim.x) ss += x[c] * x[c];
  }
  ss = block_sum_float(ss, red);
  const float nrm = fmaxf(sqrtf(ss), 1e-12f);
  if (vec) {
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(x + c);
      *reinterpret_cast<float4*>(y + c) = make_float4(v.x / nrm, v.y / nrm, v.z / nrm, v.w / nrm);
    }
  } else {
    for (int c = threadIdx.x; c < d; c += blockDim.x) y[c] = x[c] / nrm;
  }
  float* no = second ? n1 : n0;
  if (no && threadIdx.x == 0) no[r] = nrm;
}

// Both normalize backward passes of a loss in ONE launch; the second upstream gradient is the sum of up to three
// separately computed products (dy1 = a + b + c, b / c optional), added on the fly.
__global__ void __launch_bounds__(256) normalize_bwd2_kernel(const float* __restrict__ dy0, const float* __restrict__ y0,
                                                              const float* __restrict__ nrm0, float* __restrict__ dx0,
                                                              const float* __restrict__ dy1a, const float* __restrict__ dy1b,
                                                              const float* __restrict__ dy1c, const float* __restrict__ y1,
                                                              const float* __restrict__ nrm1, float* __restrict__ dx1,
                                                              int rows, int d) {
  __shared__ float red[8];
  const bool second = (int)blockIdx.x >= rows;
  const int r = second ? (int)blockIdx.x - rows : (int)blockIdx.x;
  const int64_t o = (int64_t)r * d;
  const float* a = (second ? dy1a : dy0) + o;
  const float* b = second && dy1b ? dy1b + o : nullptr;
  const float* c3 = second && dy1c ? dy1c + o : nullptr;
  const float* y = (second ? y1 : y0) + o;
  float* dx = (second ? dx1 : dx0) + o;
  auto up = [&](int c) { return a[c] + (b ? b[c] : 0.f) + (c3 ? c3[c] : 0.f); };
  float dot = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) dot += up(c) * y[c];
  dot = block_sum_float(dot, red);
  const float n = (second ? nrm1 : nrm0)[r];
  for (int c = threadIdx.x; c < d; c += blockDim.x) dx[c] = (up(c) - y[c] * dot) / n;
}

// Row i of the cross-modal soft-max (loss_ctrs.py:166-177), in place:
//   inter[i,:] (already / T)  -> d loss / d inter[i,:]
//   ori[i,:]   (raw dots)     -> d loss / d ori[i,:]   (mask and 1/T folded in)
__global__ void __launch_bounds__(256) contrastive_row_kernel(float* __restrict__ inter, float* __restrict__ ori, int b,
                                                               int n_keys, int mask_col0, int no_intra, float inv_t,
                                                               float neg_w, float scale, const float* __restrict__ weight,
                                                               float* __restrict__ partial) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  float* irow = inter + (int64_t)i * b;
  float* orow = ori + (int64_t)i * n_keys;
  const int mcol = mask_col0 + i;
  float se = 0.f, sq = 0.f;
  for (int j = threadIdx.x; j < b; j += blockDim.x) se += expf(irow[j]);
  for (int q = threadIdx.x; q < n_keys; q += blockDim.x) {
    const float logit = (no_intra || q == mcol) ? 0.f : orow[q] * inv_t;   // masked logit is 0 -> exp = 1
    sq += expf(logit);
  }
  se = block_sum_float(se, red);
  sq = block_sum_float(sq, red);
  const float z = se + neg_w * sq;
  const float w = weight[i];
  const float dii = irow[i];
  __syncthreads();
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    float g = scale * w * (expf(irow[j]) / z);
    if (j == i) g -= scale * w;
    irow[j] = g;
  }
  for (int q = threadIdx.x; q < n_keys; q += blockDim.x) {
    float g = 0.f;
    if (!no_intra && q != mcol) g = scale * w * neg_w * expf(orow[q] * inv_t) / z * inv_t;
    orow[q] = g;
  }
  if (threadIdx.x == 0) partial[i] = -logf(expf(dii) / z) * w;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ---------------------------------------------------------------------------------------------
// GEMM front-end of the loss tiles.  Tensor-core path: 3xTF32 (operands split into tf32 hi + lo parts by
// launch_split_*, K-concatenated, ONE tcgen05 tf32 GEMM with fp32 accumulation -> fp32-grade, ~1e-6 relative),
// with K split over idle SMs when the output is only a few tiles (the B x B tile).  Fallback (odd sizes): the
// fp32 FMA kernel above.
//   C[m, n] = alpha * sum_k X[m, k] * Y[n, k]     X given as rows [m, k] (xt = false) or transposed [k, m] (xt = true)
// ---------------------------------------------------------------------------------------------
struct TcScratch {
  float* xa;       // [m, 3k]
  float* yb;       // [n, 3k]
  void* ksplit;    // K-split partial tiles
  size_t ksplit_bytes;
};

static bool tc_ok(int m, int n, int k) { return k % 4 == 0 && m >= 32 && n >= 32; }

static Gemm3xDesc g3_desc(const float* x, bool xt, int64_t ldx, const float* y, bool yt, int64_t ldy, float* c, int64_t ldc,
                          int m, int n, int k, float alpha) {
  Gemm3xDesc g{};
  g.a = x; g.lda = ldx; g.a_mn = xt ? 1 : 0;
  g.b = y; g.ldb = ldy; g.b_mn = yt ? 1 : 0;
  g.c = c; g.ldc = ldc; g.m = m; g.n = n; g.k = k; g.alpha = alpha;
  return g;
}

static const bool g_loss_unfused = getenv("FRX_LOSS_UNFUSED") != nullptr;   // A/B switch: the round-1 split + GEMM chain

// One or two independent products in ONE launch of the in-kernel-split GEMM (+ one reduction launch per product when the
// outputs are so few tiles that K is split over the SMs).  Returns FRX_E_UNSUPPORTED when an operand layout rules it out.
static int gemm3x_run(cudaStream_t st, const TcScratch& ws, const Gemm3xDesc* g, int count) {
  for (int i = 0; i < count; ++i)
    if (!gemm3x_supported(g[i])) return FRX_E_UNSUPPORTED;
  int ksplit = gemm3x_plan_ksplit(g, count);
  const size_t avail = ws.ksplit_bytes / sizeof(float);
  while (ksplit > 1 && gemm3x_partial_floats(g, count, ksplit) > avail) --ksplit;
  return gemm3x_launch(st, g, count, ksplit, false, reinterpret_cast<float*>(ws.ksplit), avail);
}

static int gemm_nt(cudaStream_t st, bool tc, const TcScratch& ws, const float* x, bool xt, int64_t ldx, const float* y, bool yt,
                   int64_t ldy, float* c, int64_t ldc, int m, int n, int k, float alpha) {
  if (tc && !g_loss_unfused) {
    const Gemm3xDesc g = g3_desc(x, xt, ldx, y, yt, ldy, c, ldc, m, n, k, alpha);
    const int rc = gemm3x_run(st, ws, &g, 1);
    if (rc != FRX_E_UNSUPPORTED) return rc;
  }
  if (tc) {
    if (xt) launch_split_transpose(x, k, m, ldx, 1.0f, 0, ws.xa, st); else launch_split_rows(x, nullptr, m, k, ldx, 1.0f, 0, ws.xa, st);
    if (yt) launch_split_transpose(y, k, n, ldy, 1.0f, 1, ws.yb, st); else launch_split_rows(y, nullptr, n, k, ldy, 1.0f, 1, ws.yb, st);
    return dense_tf32_scaled(ws.xa, 3 * (int64_t)k, ws.yb, 3 * (int64_t)k, m, n, 3 * k, c, ldc, alpha, st, ws.ksplit,
                             ws.ksplit_bytes);
  }
  // X(m,k): rows -> (ldx, 1), transposed -> (1, ldx);   Y as B(k,n): rows [n,k] -> (sbk=1, sbn=ldy), transposed [k,n] -> (ldy, 1)
  sgemm(st, x, xt ? 1 : ldx, xt ? ldx : 1, y, yt ? ldy : 1, yt ? 1 : ldy, c, ldc, m, n, k, alpha, 0.f);
  return FRX_OK;
}

// scratch of one loss op with batch b, width d and nk key rows (nk = b without a queue):
//   xa: split A operand, at most [b, 3 * max(d, nk)]      yb: split B operand, at most 3 * d * max(b, nk) floats
//   ks: K-split partial tiles of the B x B GEMMs (32 splits worth)
static size_t xa_bytes(int b, int d, int nk) { return align256((size_t)b * 3 * (size_t)(d > nk ? d : nk) * sizeof(float)); }
static size_t yb_bytes(int b, int d, int nk) { return align256((size_t)3 * d * (size_t)(b > nk ? b : nk) * sizeof(float)); }
// K-split partial tiles of the b x b products, or the shared-tile slots of the stream mapping (2 x 64 KB per SM)
static size_t ks_bytes(int b) {
  const size_t split = (size_t)b * b * sizeof(float) * 32, stream = (size_t)num_sms() * 2 * 128 * 128 * sizeof(float);
  return align256(split > stream ? split : stream);
}
static size_t tc_scratch_bytes(int b, int d, int nk) { return xa_bytes(b, d, nk) + yb_bytes(b, d, nk) + ks_bytes(b); }

// ---------------------------------------------------------------------------------------------
// A14: CrossCLR_onlyIntraModality (loss_ctrs.py:52-117) and LabLoss (loss.py:55-63) on the same tile machinery.
// ---------------------------------------------------------------------------------------------
// Row i of both soft-maxes of CrossCLR.  Logits (already / T):  lbp[i,j] = bn_i.pn_j,  lbb = bn.bn^T,  lpp = pn.pn^T.
//   brand side: [ lbp[i,:] , w * lbb[i,:] with the diagonal entry replaced by 0 ]      weight rank_b[i]
//   post side:  [ lbp[:,i] , w * lpp[i,:] with the diagonal entry replaced by 0 ]      weight rank_p[i]
// (the reference multiplies the intra logits by an off-diagonal mask, loss_ctrs.py:97-99, so the diagonal stays
// in the soft-max as a logit of 0).  Writes d loss / d logits:
//   g1[i,:]  brand side w.r.t. lbp[i,:]        g2[i,:]  post side w.r.t. lbp[:,i]  (i.e. the TRANSPOSED position)
//   lbb[i,:] and lpp[i,:] are overwritten by their own gradients.
__global__ void __launch_bounds__(256) crossclr_row_kernel(const float* __restrict__ lbp, float* __restrict__ lbb,
                                                            float* __restrict__ lpp, int b, float neg_w, float scale,
                                                            const float* __restrict__ rank_p, const float* __restrict__ rank_b,
                                                            float* __restrict__ g1, float* __restrict__ g2,
                                                            float* __restrict__ partial) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float* brow = lbp + (int64_t)i * b;
  float* bb = lbb + (int64_t)i * b;
  float* pp = lpp + (int64_t)i * b;
  // soft-max maxima (F.softmax subtracts the row maximum)
  float mb = 0.f, mp = 0.f;                       // the zeroed diagonal intra logit is always present
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    mb = fmaxf(mb, brow[j]);
    mp = fmaxf(mp, lbp[(int64_t)j * b + i]);
    if (j != i) { mb = fmaxf(mb, neg_w * bb[j]); mp = fmaxf(mp, neg_w * pp[j]); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    mp = fmaxf(mp, __shfl_xor_sync(0xffffffffu, mp, o));
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mb;
  __syncthreads();
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mb = fmaxf(mb, red[w]);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mp;
  __syncthreads();
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mp = fmaxf(mp, red[w]);
  float zb = 0.f, zp = 0.f;
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    zb += expf(brow[j] - mb) + expf((j == i ? 0.f : neg_w * bb[j]) - mb);
    zp += expf(lbp[(int64_t)j * b + i] - mp) + expf((j == i ? 0.f : neg_w * pp[j]) - mp);
  }
  zb = block_sum_float(zb, red);
  zp = block_sum_float(zp, red);
  const float wb = rank_b[i], wp = rank_p[i];
  const float dii = brow[i];
  __syncthreads();
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    const float sb = expf(brow[j] - mb) / zb, sp = expf(lbp[(int64_t)j * b + i] - mp) / zp;
    g1[(int64_t)i * b + j] = scale * wb * (sb - (j == i ? 1.f : 0.f));
    g2[(int64_t)i * b + j] = scale * wp * (sp - (j == i ? 1.f : 0.f));
    const float nb_ = bb[j], np_ = pp[j];
    bb[j] = j == i ? 0.f : scale * wb * neg_w * expf(neg_w * nb_ - mb) / zb;
    pp[j] = j == i ? 0.f : scale * wp * neg_w * expf(neg_w * np_ - mp) / zp;
  }
  if (threadIdx.x == 0) partial[i] = wb * (mb + logf(zb) - dii) + wp * (mp + logf(zp) - dii);
}

// xb[i, :] = [ g1[i,j] + g2[j,i] | gbb[i,j] + gbb[j,i] ]     (d loss / d (bn.pn^T) and the symmetrised intra gradient)
// xp[i, :] = [ g1[j,i] + g2[i,j] | gpp[i,j] + gpp[j,i] ]     both [b, 2b] row-major
__global__ void __launch_bounds__(256) crossclr_combine_kernel(const float* __restrict__ g1, const float* __restrict__ g2,
                                                                const float* __restrict__ gbb, const float* __restrict__ gpp,
                                                                int b, float* __restrict__ xb, float* __restrict__ xp) {
  __shared__ float t1[32][33], t2[32][33], tb[32][33], tp[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 32; r += 8) {                 // transposed block (rows j0.., cols i0..)
    const int jj = j0 + ty + r, ii = i0 + tx;
    const bool ok = jj < b && ii < b;
    t1[ty + r][tx] = ok ? g1[(int64_t)jj * b + ii] : 0.f;
    t2[ty + r][tx] = ok ? g2[(int64_t)jj * b + ii] : 0.f;
    tb[ty + r][tx] = ok ? gbb[(int64_t)jj * b + ii] : 0.f;
    tp[ty + r][tx] = ok ? gpp[(int64_t)jj * b + ii] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int ii = i0 + ty + r, jj = j0 + tx;
    if (ii < b && jj < b) {
      const int64_t src = (int64_t)ii * b + jj, dst = (int64_t)ii * 2 * b + jj;
      xb[dst] = g1[src] + t2[tx][ty + r];
      xb[dst + b] = gbb[src] + tb[tx][ty + r];
      xp[dst] = t1[tx][ty + r] + g2[src];
      xp[dst + b] = gpp[src] + tp[tx][ty + r];
    }
  }
}

// LabLoss row i: s[i,:] = cos(brand_i, brand_:) -> partial[i] = sum_j exp(s_ij with the diagonal set to 0);
// s[i,j] <- d loss / d s[i,j] = exp(s_ij) / b off the diagonal, 0 on it (masked_fill, loss.py:59-60).
__global__ void __launch_bounds__(256) lab_row_kernel(float* __restrict__ s, int b, float* __restrict__ partial) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  float* row = s + (int64_t)i * b;
  const float inv_b = 1.0f / (float)b;
  float acc = 0.f;
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    const float e = j == i ? 1.0f : expf(row[j]);
    acc += e;
    row[j] = j == i ? 0.f : e * inv_b;
  }
  acc = block_sum_float(acc, red);
  if (threadIdx.x == 0) partial[i] = acc;
}

// out = (sum(partial) - offset) * scale
__global__ void __launch_bounds__(256) reduce_offset_kernel(const float* __restrict__ partial, int n, double offset, double scale,
                                                             float* __restrict__ out) {
  __shared__ double red[8];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[0] = (float)((t - offset) * scale);
  }
}

// x / ||x|| without the eps clamp (loss.l2norm, loss.py:20-24); norm_out optional
__global__ void __launch_bounds__(256) l2norm_rows_kernel(const float* __restrict__ x, int d, float* __restrict__ y,
                                                           float* __restrict__ norm_out) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  float ss = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) { const float v = x[(int64_t)r * d + c]; ss += v * v; }
  ss = block_sum_float(ss, red);
  const float nrm = sqrtf(ss);
  for (int c = threadIdx.x; c < d; c += blockDim.x) y[(int64_t)r * d + c] = x[(int64_t)r * d + c] / nrm;
  if (norm_out && threadIdx.x == 0) norm_out[r] = nrm;
}

// ---------------------------------------------------------------------------------------------
// Opt-in: the hardest-negative ("max of violations", VSE++) hinge that north_star names and loss.py's ignored
// `max_violation` flag stands for:  loss = sum_i max_j cost_p[i,j] + sum_j max_i cost_b[i,j]  with
// cost_p = [m + S - S_ii]_+, cost_b = [m + S - S_jj]_+, same-brand entries zeroed as in loss.py:116-119, no rank weights.
// Block i takes row i AND column i of the tile: the two maxima, their first positions (-1 when the maximum is 0: no
// gradient), partial[i] = both maxima.
__global__ void __launch_bounds__(256) vsepp_select_kernel(const float* __restrict__ s, const int64_t* __restrict__ ids, int b,
                                                            float margin, int* __restrict__ row_arg, int* __restrict__ col_arg,
                                                            float* __restrict__ partial) {
  __shared__ float vmax[8];
  __shared__ int vidx[8];
  const int i = blockIdx.x;
  const float dii = s[(int64_t)i * b + i];
  const int64_t idi = ids[i];
  float best[2] = {0.f, 0.f};
  int arg[2] = {-1, -1};
  for (int j = threadIdx.x; j < b; j += blockDim.x) {
    if (ids[j] == idi) continue;
    const float cp = margin + s[(int64_t)i * b + j] - dii;       // row i, negative brand j
    const float cb = margin + s[(int64_t)j * b + i] - dii;       // column i, negative post j
    if (cp > best[0]) { best[0] = cp; arg[0] = j; }
    if (cb > best[1]) { best[1] = cb; arg[1] = j; }
  }
  float total = 0.f;
  for (int w = 0; w < 2; ++w) {
    float v = best[w]; int a = arg[w];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oa = __shfl_xor_sync(0xffffffffu, a, o);
      if (ov > v || (ov == v && oa >= 0 && (a < 0 || oa < a))) { v = ov; a = oa; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { vmax[threadIdx.x >> 5] = v; vidx[threadIdx.x >> 5] = a; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
        if (vmax[k] > v || (vmax[k] == v && vidx[k] >= 0 && (a < 0 || vidx[k] < a))) { v = vmax[k]; a = vidx[k]; }
      (w == 0 ? row_arg : col_arg)[i] = v > 0.f ? a : -1;
      total += v > 0.f ? v : 0.f;
    }
  }
  if (threadIdx.x == 0) partial[i] = total;
}
// dS (zeroed before): +scale at the selected entries, -scale on the diagonal for each of them.  Sums of equal
// magnitudes: exact in fp32, so the atomics' order does not matter.
__global__ void __launch_bounds__(256) vsepp_scatter_kernel(const int* __restrict__ row_arg, const int* __restrict__ col_arg, int b,
                                                             float scale, float* __restrict__ ds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  const int jr = row_arg[i], ic = col_arg[i];
  if (jr >= 0) { atomicAdd(ds + (int64_t)i * b + jr, scale); atomicAdd(ds + (int64_t)i * b + i, -scale); }
  if (ic >= 0) { atomicAdd(ds + (int64_t)ic * b + i, scale); atomicAdd(ds + (int64_t)i * b + i, -scale); }
}

}  // namespace frx

extern "C" {

size_t frx_triplet_workspace_bytes(int b, int d) {
  if (b <= 0 || d <= 0) return 0;
  return 2 * frx::align256((size_t)b * b * 4) + 4 * frx::align256((size_t)b * 4) + frx::tc_scratch_bytes(b, d, b) + 256;
}

int frx_triplet_fwd_bwd(const int64_t* brand_ids, const float* brand, const float* post, int b, int d, float margin,
                        int mean_style, float* loss, float* d_brand, float* d_post, void* workspace,
                        size_t workspace_bytes, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(brand_ids && brand && post && loss, "frx_triplet_fwd_bwd: NULL pointer");
  FRX_CHECK_ARG(b > 0 && d > 0, "frx_triplet_fwd_bwd: bad sizes");
  FRX_CHECK_ARG((d_brand == nullptr) == (d_post == nullptr), "frx_triplet_fwd_bwd: gradients go together");
  if (!workspace || workspace_bytes < frx_triplet_workspace_bytes(b, d)) {
    set_error("frx_triplet_fwd_bwd: workspace %zu bytes, need %zu", workspace_bytes, frx_triplet_workspace_bytes(b, d));
    return FRX_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  float* s = reinterpret_cast<float*>(w); w += align256((size_t)b * b * 4);
  float* ds = reinterpret_cast<float*>(w); w += align256((size_t)b * b * 4);
  float* rank_p = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  float* rank_b = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  float* diag = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  float* partial = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  TcScratch ts;
  ts.xa = reinterpret_cast<float*>(w); w += xa_bytes(b, d, b);
  ts.yb = reinterpret_cast<float*>(w); w += yb_bytes(b, d, b);
  ts.ksplit = w;
  ts.ksplit_bytes = ks_bytes(b);
  const bool tc = tc_ok(b, b, d) && b % 4 == 0;
  const float scale = mean_style ? 1.0f / ((float)b * (float)b) : 1.0f;
  // ---- fused path: 3 launches.  Tile GEMM with the 3xTF32 split inside the kernel and K split over the SMs (raw partial
  // tiles) -> one cooperative kernel for the whole tile stage -> both gradient GEMMs in one grid (dS and the embeddings
  // read in place, as K-major or transposed operands).
  {
    Gemm3xDesc gs{};
    gs.a = post; gs.lda = d; gs.a_mn = 0; gs.b = brand; gs.ldb = d; gs.b_mn = 0;
    gs.c = s; gs.ldc = b; gs.m = b; gs.n = b; gs.k = d; gs.alpha = 1.f;
    Gemm3xDesc gg[2]{};
    gg[0].a = ds; gg[0].lda = b; gg[0].a_mn = 0; gg[0].b = brand; gg[0].ldb = d; gg[0].b_mn = 1;     // dPost  = dS   . brand
    gg[0].c = d_post; gg[0].ldc = d; gg[0].m = b; gg[0].n = d; gg[0].k = b; gg[0].alpha = 1.f;
    gg[1].a = ds; gg[1].lda = b; gg[1].a_mn = 1; gg[1].b = post; gg[1].ldb = d; gg[1].b_mn = 1;      // dBrand = dS^T . post
    gg[1].c = d_brand; gg[1].ldc = d; gg[1].m = b; gg[1].n = d; gg[1].k = b; gg[1].alpha = 1.f;
    const size_t tile_smem = (size_t)b * sizeof(float);
    if (!g_loss_unfused && gemm3x_supported(gs) && gemm3x_supported(gg[0]) && gemm3x_supported(gg[1]) && tile_smem <= 40 * 1024 &&
        triplet_tile_fits(b, tile_smem)) {
      int ksplit = gemm3x_plan_ksplit(&gs, 1);
      const size_t avail = ts.ksplit_bytes / sizeof(float);
      while (ksplit > 1 && gemm3x_partial_floats(&gs, 1, ksplit) > avail) --ksplit;
      int rc = gemm3x_launch(st, &gs, 1, ksplit, true, reinterpret_cast<float*>(ts.ksplit), avail);
      if (rc) return rc;
      const float* part = reinterpret_cast<const float*>(ts.ksplit);
      float* ds_arg = d_post ? ds : nullptr;
      void* args[] = {(void*)&part, (void*)&ksplit, (void*)&b, (void*)&brand_ids, (void*)&margin, (void*)&scale, (void*)&s,
                      (void*)&rank_p, (void*)&rank_b, (void*)&diag, (void*)&ds_arg, (void*)&partial, (void*)&loss};
      FRX_CUDA(cudaLaunchCooperativeKernel((const void*)triplet_tile_kernel, dim3(b), dim3(256), args, tile_smem, st));
      if (d_post) {
        rc = gemm3x_launch(st, gg, 2, 1, false, reinterpret_cast<float*>(ts.ksplit), avail);   // the tile stage has consumed the partial tiles
        if (rc) return rc;
      }
      return FRX_OK;
    }
  }
  // ---- general path (odd sizes, very large batches)
  // S[i,j] = post_i . brand_j   (loss.py:91-93)
  int rc = gemm_nt(st, tc, ts, post, false, d, brand, false, d, s, b, b, b, d, 1.f);
  if (rc) return rc;
  tile_rank_kernel<<<b, 128, 0, st>>>(s, b, rank_p, rank_b, diag);
  triplet_row_kernel<<<b, 256, 0, st>>>(s, brand_ids, b, margin, scale, rank_p, rank_b, diag, ds, partial);
  reduce_partials_kernel<<<1, 256, 0, st>>>(partial, b, scale, loss);
  if (d_post) {
    rc = gemm_nt(st, tc, ts, ds, false, b, brand, true, d, d_post, d, b, d, b, 1.f);     // dPost  = dS   . brand
    if (rc) return rc;
    rc = gemm_nt(st, tc, ts, ds, true, b, post, true, d, d_brand, d, b, d, b, 1.f);      // dBrand = dS^T . post
    if (rc) return rc;
  }
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_vsepp_fwd_bwd(const int64_t* brand_ids, const float* brand, const float* post, int b, int d, float margin,
                      int mean_style, float* loss, float* d_brand, float* d_post, void* workspace, size_t workspace_bytes,
                      void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(brand_ids && brand && post && loss, "frx_vsepp_fwd_bwd: NULL pointer");
  FRX_CHECK_ARG(b > 0 && d > 0, "frx_vsepp_fwd_bwd: bad sizes");
  FRX_CHECK_ARG((d_brand == nullptr) == (d_post == nullptr), "frx_vsepp_fwd_bwd: gradients go together");
  if (!workspace || workspace_bytes < frx_triplet_workspace_bytes(b, d)) {
    set_error("frx_vsepp_fwd_bwd: workspace %zu bytes, need %zu", workspace_bytes, frx_triplet_workspace_bytes(b, d));
    return FRX_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  float* s = reinterpret_cast<float*>(w); w += align256((size_t)b * b * 4);
  float* ds = reinterpret_cast<float*>(w); w += align256((size_t)b * b * 4);
  int* row_arg = reinterpret_cast<int*>(w); w += align256((size_t)b * 4);
  int* col_arg = reinterpret_cast<int*>(w); w += align256((size_t)b * 4);
  w += align256((size_t)b * 4);
  float* partial = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  TcScratch ts;
  ts.xa = reinterpret_cast<float*>(w); w += xa_bytes(b, d, b);
  ts.yb = reinterpret_cast<float*>(w); w += yb_bytes(b, d, b);
  ts.ksplit = w;
  ts.ksplit_bytes = ks_bytes(b);
  const bool tc = tc_ok(b, b, d) && b % 4 == 0;
  const float scale = mean_style ? 1.0f / (float)b : 1.0f;
  int rc = gemm_nt(st, tc, ts, post, false, d, brand, false, d, s, b, b, b, d, 1.f);      // S[i,j] = post_i . brand_j
  if (rc) return rc;
  vsepp_select_kernel<<<b, 256, 0, st>>>(s, brand_ids, b, margin, row_arg, col_arg, partial);
  reduce_partials_kernel<<<1, 256, 0, st>>>(partial, b, scale, loss);
  if (d_post) {
    FRX_CUDA(cudaMemsetAsync(ds, 0, (size_t)b * b * sizeof(float), st));
    vsepp_scatter_kernel<<<(b + 255) / 256, 256, 0, st>>>(row_arg, col_arg, b, scale, ds);
    const Gemm3xDesc gg[2] = {g3_desc(ds, false, b, brand, true, d, d_post, d, b, d, b, 1.f),     // dPost  = dS   . brand
                              g3_desc(ds, true, b, post, true, d, d_brand, d, b, d, b, 1.f)};     // dBrand = dS^T . post
    rc = (tc && !g_loss_unfused) ? gemm3x_run(st, ts, gg, 2) : FRX_E_UNSUPPORTED;
    if (rc == FRX_E_UNSUPPORTED) {
      rc = gemm_nt(st, tc, ts, ds, false, b, brand, true, d, d_post, d, b, d, b, 1.f);
      if (rc) return rc;
      rc = gemm_nt(st, tc, ts, ds, true, b, post, true, d, d_brand, d, b, d, b, 1.f);
    }
    if (rc) return rc;
  }
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

size_t frx_contrastive_workspace_bytes(int b, int d, int n_keys) {
  if (b <= 0 || d <= 0) return 0;
  const int nk = n_keys > 0 ? n_keys : b;
  return frx::align256((size_t)b * b * 4) + frx::align256((size_t)b * nk * 4) + 7 * frx::align256((size_t)b * d * 4) +
         6 * frx::align256((size_t)b * 4) + frx::tc_scratch_bytes(b, d, nk) + 256;
}

int frx_contrastive_fwd_bwd(const float* brand, const float* post, int b, int d, const float* keys, int n_keys,
                            int mask_col0, int no_intra, float temperature, float negative_weight, int mean_style,
                            float* loss, float* d_brand, float* d_post, void* workspace, size_t workspace_bytes,
                            void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(brand && post && loss, "frx_contrastive_fwd_bwd: NULL pointer");
  FRX_CHECK_ARG(b > 0 && d > 0, "frx_contrastive_fwd_bwd: bad sizes");
  FRX_CHECK_ARG((keys == nullptr) == (n_keys == 0), "frx_contrastive_fwd_bwd: keys and n_keys disagree");
  FRX_CHECK_ARG((d_brand == nullptr) == (d_post == nullptr), "frx_contrastive_fwd_bwd: gradients go together");
  const int nk = keys ? n_keys : b;
  FRX_CHECK_ARG(mask_col0 >= 0 && mask_col0 + b <= nk, "frx_contrastive_fwd_bwd: mask columns %d..%d outside %d keys",
                mask_col0, mask_col0 + b - 1, nk);
  const size_t need = frx_contrastive_workspace_bytes(b, d, n_keys);
  if (!workspace || workspace_bytes < need) {
    set_error("frx_contrastive_fwd_bwd: workspace %zu bytes, need %zu", workspace_bytes, need);
    return FRX_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  float* inter = reinterpret_cast<float*>(w); w += align256((size_t)b * b * 4);
  float* ori = reinterpret_cast<float*>(w); w += align256((size_t)b * nk * 4);
  float* bn = reinterpret_cast<float*>(w); w += align256((size_t)b * d * 4);
  float* pn = reinterpret_cast<float*>(w); w += align256((size_t)b * d * 4);
  float* dbn = reinterpret_cast<float*>(w); w += align256((size_t)b * d * 4);
  w += align256((size_t)b * d * 4);                              // (was: the summed d_pn; now formed inside normalize_bwd2)
  float* t1 = reinterpret_cast<float*>(w); w += align256((size_t)b * d * 4);
  float* t2 = reinterpret_cast<float*>(w); w += align256((size_t)b * d * 4);
  float* t3 = reinterpret_cast<float*>(w); w += align256((size_t)b * d * 4);
  float* weight = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  float* rank_b = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  float* diag = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  float* partial = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  float* nrm_b = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  float* nrm_p = reinterpret_cast<float*>(w); w += align256((size_t)b * 4);
  TcScratch ts;
  ts.xa = reinterpret_cast<float*>(w); w += xa_bytes(b, d, nk);
  ts.yb = reinterpret_cast<float*>(w); w += yb_bytes(b, d, nk);
  ts.ksplit = w;
  ts.ksplit_bytes = ks_bytes(b);
  const bool tc = tc_ok(b, b, d) && b % 4 == 0 && nk % 4 == 0;
  const float inv_t = 1.0f / temperature;
  const float scale = mean_style ? 1.0f / (float)b : 1.0f;
  normalize_rows2_kernel<<<2 * b, 256, 0, st>>>(brand, post, b, d, bn, pn, nrm_b, nrm_p);
  const float* kk = keys ? keys : pn;
  // The RAW tile post . brand^T for the rank weights (loss_ctrs.py:182-192; parked in `ori`, which is written only after
  // tile_rank has read it) and inter[i,j] = bn_i . pn_j / T: both B x B products in one grid.
  float* raw_tile = ori;
  const Gemm3xDesc gt[2] = {g3_desc(post, false, d, brand, false, d, raw_tile, b, b, b, d, 1.f),
                            g3_desc(bn, false, d, pn, false, d, inter, b, b, b, d, inv_t)};
  int rc = (tc && !g_loss_unfused) ? gemm3x_run(st, ts, gt, 2) : FRX_E_UNSUPPORTED;
  if (rc == FRX_E_UNSUPPORTED) {
    rc = gemm_nt(st, tc, ts, post, false, d, brand, false, d, raw_tile, b, b, b, d, 1.f);
    if (rc) return rc;
    rc = gemm_nt(st, tc, ts, bn, false, d, pn, false, d, inter, b, b, b, d, inv_t);
  }
  if (rc) return rc;
  tile_rank_kernel<<<b, 128, 0, st>>>(raw_tile, b, weight, rank_b, diag);
  // ori[i,q] = pn_i . key_q
  rc = gemm_nt(st, tc, ts, pn, false, d, kk, false, d, ori, nk, b, nk, d, 1.f);
  if (rc) return rc;
  contrastive_row_kernel<<<b, 256, 0, st>>>(inter, ori, b, nk, mask_col0, no_intra, inv_t, negative_weight, scale,
                                            weight, partial);
  reduce_partials_kernel<<<1, 256, 0, st>>>(partial, b, scale, loss);
  if (d_post) {
    // d_bn = dInter . pn / T  and  t1 = dInter^T . bn / T : one grid
    const Gemm3xDesc gp[2] = {g3_desc(inter, false, b, pn, true, d, dbn, d, b, d, b, inv_t),
                              g3_desc(inter, true, b, bn, true, d, t1, d, b, d, b, inv_t)};
    rc = (tc && !g_loss_unfused) ? gemm3x_run(st, ts, gp, 2) : FRX_E_UNSUPPORTED;
    if (rc == FRX_E_UNSUPPORTED) {
      rc = gemm_nt(st, tc, ts, inter, false, b, pn, true, d, dbn, d, b, d, b, inv_t);
      if (rc) return rc;
      rc = gemm_nt(st, tc, ts, inter, true, b, bn, true, d, t1, d, b, d, b, inv_t);
    }
    if (rc) return rc;
    const float* sum2 = nullptr;
    const float* sum3 = nullptr;
    if (!no_intra) {
      rc = gemm_nt(st, tc, ts, ori, false, nk, kk, true, d, t2, d, b, d, nk, 1.f);         // t2 = G . keys
      if (rc) return rc;
      sum2 = t2;
      if (!keys) {                                                                         // keys = pn carry grad too
        rc = gemm_nt(st, tc, ts, ori, true, nk, pn, true, d, t3, d, b, d, b, 1.f);         // t3 = G^T . pn
        if (rc) return rc;
        sum3 = t3;
      }
    }
    // d_pn = t1 + t2 + t3 is formed inside the backward of the normalisation
    normalize_bwd2_kernel<<<2 * b, 256, 0, st>>>(dbn, bn, nrm_b, d_brand, t1, sum2, sum3, pn, nrm_p, d_post, b, d);
  }
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

size_t frx_crossclr_workspace_bytes(int b, int d) {
  if (b <= 0 || d <= 0) return 0;
  return 5 * frx::align256((size_t)b * b * 4) + 2 * frx::align256((size_t)b * 2 * b * 4) + 6 * frx::align256((size_t)b * d * 4) +
         6 * frx::align256((size_t)b * 4) + frx::tc_scratch_bytes(b, d, 2 * b) + 256;
}

int frx_crossclr_fwd_bwd(const float* brand, const float* post, int b, int d, float temperature, float negative_weight,
                         int mean_style, float* loss, float* d_brand, float* d_post, void* workspace,
                         size_t workspace_bytes, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(brand && post && loss, "frx_crossclr_fwd_bwd: NULL pointer");
  FRX_CHECK_ARG(b > 0 && d > 0, "frx_crossclr_fwd_bwd: bad sizes");
  FRX_CHECK_ARG((d_brand == nullptr) == (d_post == nullptr), "frx_crossclr_fwd_bwd: gradients go together");
  const size_t need = frx_crossclr_workspace_bytes(b, d);
  if (!workspace || workspace_bytes < need) {
    set_error("frx_crossclr_fwd_bwd: workspace %zu bytes, need %zu", workspace_bytes, need);
    return FRX_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) { float* p = reinterpret_cast<float*>(w); w += align256(bytes); return p; };
  const size_t bb4 = (size_t)b * b * 4, bd4 = (size_t)b * d * 4;
  float* lbp = take(bb4); float* lbb = take(bb4); float* lpp = take(bb4); float* g1 = take(bb4); float* g2 = take(bb4);
  float* xb = take(2 * bb4); float* xp = take(2 * bb4);
  float* n1 = take(2 * bd4);      // [pn ; bn]   (2b x d)
  float* n2 = take(2 * bd4);      // [bn ; pn]
  float* dbn = take(bd4); float* dpn = take(bd4);
  float* rank_p = take((size_t)b * 4); float* rank_b = take((size_t)b * 4); float* diag = take((size_t)b * 4);
  float* partial = take((size_t)b * 4); float* nrm_b = take((size_t)b * 4); float* nrm_p = take((size_t)b * 4);
  TcScratch ts;
  ts.xa = reinterpret_cast<float*>(w); w += xa_bytes(b, d, 2 * b);
  ts.yb = reinterpret_cast<float*>(w); w += yb_bytes(b, d, 2 * b);
  ts.ksplit = w;
  ts.ksplit_bytes = ks_bytes(b);
  const bool tc = tc_ok(b, b, d) && b % 4 == 0;
  const float inv_t = 1.0f / temperature;
  const float scale = mean_style ? 0.5f / (float)b : 0.5f;
  float* pn = n1; float* bn = n1 + (size_t)b * d;
  normalize_rows2_kernel<<<2 * b, 256, 0, st>>>(post, brand, b, d, pn, bn, nrm_p, nrm_b);
  FRX_CUDA(cudaMemcpyAsync(n2, bn, bd4, cudaMemcpyDeviceToDevice, st));
  FRX_CUDA(cudaMemcpyAsync(n2 + (size_t)b * d, pn, bd4, cudaMemcpyDeviceToDevice, st));
  // Four B x B products, two per grid: the RAW tile scores[i][j] = post_i . brand_j for the rank weights
  // (loss_ctrs.py:62-77; parked in g1, which crossclr_row_kernel overwrites later) with bn . pn^T / T, then the two Gram tiles.
  float* raw_tile = g1;
  const Gemm3xDesc ga[2] = {g3_desc(post, false, d, brand, false, d, raw_tile, b, b, b, d, 1.f),
                            g3_desc(bn, false, d, pn, false, d, lbp, b, b, b, d, inv_t)};
  const Gemm3xDesc gb[2] = {g3_desc(bn, false, d, bn, false, d, lbb, b, b, b, d, inv_t),
                            g3_desc(pn, false, d, pn, false, d, lpp, b, b, b, d, inv_t)};
  int rc = (tc && !g_loss_unfused) ? gemm3x_run(st, ts, ga, 2) : FRX_E_UNSUPPORTED;
  if (rc == FRX_E_UNSUPPORTED) {
    rc = gemm_nt(st, tc, ts, post, false, d, brand, false, d, raw_tile, b, b, b, d, 1.f);
    if (rc) return rc;
    rc = gemm_nt(st, tc, ts, bn, false, d, pn, false, d, lbp, b, b, b, d, inv_t);
  }
  if (rc) return rc;
  tile_rank_kernel<<<b, 128, 0, st>>>(raw_tile, b, rank_p, rank_b, diag);
  rc = (tc && !g_loss_unfused) ? gemm3x_run(st, ts, gb, 2) : FRX_E_UNSUPPORTED;
  if (rc == FRX_E_UNSUPPORTED) {
    rc = gemm_nt(st, tc, ts, bn, false, d, bn, false, d, lbb, b, b, b, d, inv_t);
    if (rc) return rc;
    rc = gemm_nt(st, tc, ts, pn, false, d, pn, false, d, lpp, b, b, b, d, inv_t);
  }
  if (rc) return rc;
  crossclr_row_kernel<<<b, 256, 0, st>>>(lbp, lbb, lpp, b, negative_weight, scale, rank_p, rank_b, g1, g2, partial);
  reduce_partials_kernel<<<1, 256, 0, st>>>(partial, b, scale, loss);
  if (d_post) {
    dim3 cg((b + 31) / 32, (b + 31) / 32);
    crossclr_combine_kernel<<<cg, 256, 0, st>>>(g1, g2, lbb, lpp, b, xb, xp);
    const bool tc2 = tc && (2 * b) % 4 == 0;
    rc = gemm_nt(st, tc2, ts, xb, false, 2 * b, n1, true, d, dbn, d, b, d, 2 * b, inv_t);   // d_bn = [G_bp | G_bb+G_bb^T] . [pn ; bn] / T
    if (rc) return rc;
    rc = gemm_nt(st, tc2, ts, xp, false, 2 * b, n2, true, d, dpn, d, b, d, 2 * b, inv_t);   // d_pn = [G_bp^T | G_pp+G_pp^T] . [bn ; pn] / T
    if (rc) return rc;
    normalize_bwd2_kernel<<<2 * b, 256, 0, st>>>(dbn, bn, nrm_b, d_brand, dpn, nullptr, nullptr, pn, nrm_p, d_post, b, d);
  }
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

size_t frx_lab_workspace_bytes(int b, int d) {
  if (b <= 0 || d <= 0) return 0;
  return frx::align256((size_t)b * b * 4) + 2 * frx::align256((size_t)b * d * 4) + 2 * frx::align256((size_t)b * 4) +
         frx::tc_scratch_bytes(b, d, b) + 256;
}

int frx_lab_fwd_bwd(const float* brand, int b, int d, float* loss, float* d_brand, void* workspace, size_t workspace_bytes,
                    void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(brand && loss, "frx_lab_fwd_bwd: NULL pointer");
  FRX_CHECK_ARG(b > 0 && d > 0, "frx_lab_fwd_bwd: bad sizes");
  const size_t need = frx_lab_workspace_bytes(b, d);
  if (!workspace || workspace_bytes < need) {
    set_error("frx_lab_fwd_bwd: workspace %zu bytes, need %zu", workspace_bytes, need);
    return FRX_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) { float* p = reinterpret_cast<float*>(w); w += align256(bytes); return p; };
  float* s = take((size_t)b * b * 4);
  float* bn = take((size_t)b * d * 4); float* dbn = take((size_t)b * d * 4);
  float* partial = take((size_t)b * 4); float* nrm = take((size_t)b * 4);
  TcScratch ts;
  ts.xa = reinterpret_cast<float*>(w); w += xa_bytes(b, d, b);
  ts.yb = reinterpret_cast<float*>(w); w += yb_bytes(b, d, b);
  ts.ksplit = w;
  ts.ksplit_bytes = ks_bytes(b);
  const bool tc = tc_ok(b, b, d) && b % 4 == 0;
  l2norm_rows_kernel<<<b, 256, 0, st>>>(brand, d, bn, nrm);
  int rc = gemm_nt(st, tc, ts, bn, false, d, bn, false, d, s, b, b, b, d, 1.f);           // cosine_sim(brand, brand)
  if (rc) return rc;
  lab_row_kernel<<<b, 256, 0, st>>>(s, b, partial);
  reduce_offset_kernel<<<1, 256, 0, st>>>(partial, b, (double)b, 1.0 / (double)b, loss);   // (sum exp(s) - B) / B
  if (d_brand) {
    // s is symmetric, so is G = d loss / d s: d_bn = (G + G^T) . bn = 2 G . bn
    rc = gemm_nt(st, tc, ts, s, false, b, bn, true, d, dbn, d, b, d, b, 2.f);
    if (rc) return rc;
    normalize_bwd_kernel<<<b, 256, 0, st>>>(dbn, bn, nrm, d, d_brand);
  }
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_normalize_rows(const float* x, int rows, int d, float* out, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(x && out && rows > 0 && d > 0, "frx_normalize_rows: bad arguments");
  normalize_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, d, out, nullptr);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

}  // extern "C"
