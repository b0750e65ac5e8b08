// A6-A9: integer rank statistics behind AUC / NDCG@k / recall@k / MedR (evaluator.py:103-143,
// util/ndcg.py).  Everything here is HBM/latency-bound integer work on warp-level primitives; the
// float64 metric values are computed from these integers on the host exactly as the reference does.
#include <type_traits>
#include "common.cuh"

namespace frx {

// n_pos[b] and the best positive per brand as a packed (score, ~index) key (atomicMax).
__global__ void label_stats_kernel(const int32_t* __restrict__ labels, const float* __restrict__ pos_score,
                                   int64_t n_posts, int nb, int64_t index_base, int32_t* __restrict__ n_pos,
                                   unsigned long long* __restrict__ best_key) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_posts; j += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = labels[j];
    if (b >= 0 && b < nb) {
      atomicAdd(n_pos + b, 1);
      atomicMax(best_key + b, make_key(pos_score[j], (uint32_t)(index_base + j)));
    }
  }
}

// Same for small brand counts (nb <= kPrivBrands): the histogram and the keys are first accumulated in shared memory by
// a few large blocks (1000 brands x 1 M posts: ~7 k posts per block, shared-memory atomics with little contention), then
// flushed with one global atomic per touched brand and block -- 150 k global atomics instead of 2 M.
constexpr int kPrivBrands = 2048;
__global__ void __launch_bounds__(1024) label_stats_priv_kernel(const int32_t* __restrict__ labels,
                                                                const float* __restrict__ pos_score, int64_t n_posts, int nb,
                                                                int64_t index_base, int32_t* __restrict__ n_pos,
                                                                unsigned long long* __restrict__ best_key) {
  __shared__ int32_t s_cnt[kPrivBrands];
  __shared__ unsigned long long s_key[kPrivBrands];
  for (int b = threadIdx.x; b < nb; b += blockDim.x) { s_cnt[b] = 0; s_key[b] = 0ull; }
  __syncthreads();
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_posts; j += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = labels[j];
    if (b >= 0 && b < nb) {
      atomicAdd(&s_cnt[b], 1);
      atomicMax(&s_key[b], make_key(pos_score[j], (uint32_t)(index_base + j)));
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const int32_t c = s_cnt[b];
    if (c > 0) {
      atomicAdd(n_pos + b, c);
      atomicMax(best_key + b, s_key[b]);
    }
  }
}

__global__ void decode_best_kernel(const unsigned long long* __restrict__ best_key, const int32_t* __restrict__ n_pos,
                                   int nb, float* __restrict__ best_score, int32_t* __restrict__ best_index) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  if (n_pos[b] > 0) {
    best_score[b] = key_score(best_key[b]);
    best_index[b] = (int32_t)key_index(best_key[b]);
  } else {
    best_score[b] = -INFINITY;
    best_index[b] = -1;
  }
}

// One warp per brand: relevance bits of the first 64 ranks + rank of the first positive in the list.
__global__ void rank_from_topk_kernel(const int32_t* __restrict__ topk_index, int nb, int k,
                                      const int32_t* __restrict__ labels, int64_t n_posts, int64_t index_base,
                                      unsigned long long* __restrict__ hit_mask, int32_t* __restrict__ first_rank) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nb) return;
  unsigned long long mask = 0ull;
  int first = -1;
  for (int base = 0; base < k; base += 32) {
    const int r = base + lane;
    bool hit = false;
    if (r < k) {
      const int64_t local = (int64_t)topk_index[(size_t)b * k + r] - index_base;
      if (topk_index[(size_t)b * k + r] >= 0 && local >= 0 && local < n_posts) hit = labels[local] == b;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (base < 64) mask |= (unsigned long long)m << base;
    if (first < 0 && m) first = base + __ffs(m) - 1;
    if (first >= 0 && base >= 32) break;   // bits 0..63 and the first hit are all we need
  }
  if (lane == 0) { hit_mask[b] = mask; first_rank[b] = first; }
}

// ---- positives grouped by brand, ascending --------------------------------------------------
__global__ void seg_scan_kernel(const int32_t* __restrict__ n_pos, int nb, int64_t* __restrict__ seg_ptr,
                                unsigned long long* __restrict__ cursor) {
  // single block; nb is at most a few 10^4
  __shared__ long long carry;
  __shared__ long long wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    const int b = base + threadIdx.x;
    long long v = b < nb ? (long long)n_pos[b] : 0;
    long long incl = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      long long w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0, wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      wsum[lane] = wi - w;   // exclusive
    }
    __syncthreads();
    const long long excl = carry + wsum[warp] + incl - v;
    if (b < nb) { seg_ptr[b] = excl; cursor[b] = 0ull; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) seg_ptr[nb] = carry;
}

__global__ void scatter_pos_kernel(const int32_t* __restrict__ labels, const float* __restrict__ pos_score,
                                   int64_t n_posts, int nb, const int64_t* __restrict__ seg_ptr,
                                   unsigned long long* __restrict__ cursor, float* __restrict__ pos_sorted) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_posts; j += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = labels[j];
    if (b >= 0 && b < nb) {
      const unsigned long long slot = atomicAdd(cursor + b, 1ull);
      pos_sorted[seg_ptr[b] + (int64_t)slot] = pos_score[j];
    }
  }
}

// One block per brand: ascending bitonic sort of its positives (smem when they fit, else in place in
// global memory -- correct, slower, only for brands with > 32768 positives).
constexpr int kSortSmemFloats = 32768;
__global__ void __launch_bounds__(256) sort_segments_kernel(const int64_t* __restrict__ seg_ptr, float* __restrict__ pos_sorted) {
  extern __shared__ float sbuf[];
  const int b = blockIdx.x;
  const int64_t s0 = seg_ptr[b];
  const int64_t n = seg_ptr[b + 1] - s0;
  if (n <= 1) return;
  int64_t np2 = 1;
  while (np2 < n) np2 <<= 1;
  float* g = pos_sorted + s0;
  const bool in_smem = np2 <= kSortSmemFloats;
  if (in_smem) {
    for (int64_t i = threadIdx.x; i < np2; i += blockDim.x) sbuf[i] = i < n ? g[i] : INFINITY;
  }
  for (int64_t size = 2; size <= np2; size <<= 1) {
    for (int64_t stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int64_t t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        const int64_t lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int64_t hi = lo + stride;
        const bool asc = ((lo & size) == 0);
        if (in_smem) {
          const float a = sbuf[lo], c = sbuf[hi];
          if ((a > c) == asc) { sbuf[lo] = c; sbuf[hi] = a; }
        } else {
          // virtual +inf padding beyond n
          const float a = lo < n ? g[lo] : INFINITY, c = hi < n ? g[hi] : INFINITY;
          if ((a > c) == asc) {
            if (lo < n) g[lo] = c;
            if (hi < n) g[hi] = a;
          }
        }
      }
    }
  }
  __syncthreads();
  if (in_smem)
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) g[i] = sbuf[i];
}

// ---- exact AUC numerator + "posts before the first positive" from dense score rows ---------
// evaluator.py:111-113: sum over the brand's positives e of #{negatives el : e > el}, i.e. for every negative score s
// the number of the brand's positives above it, m - #{positives <= s}.  One block streams a row of scores once; the
// brand's sorted positives sit in shared memory behind a bucket table:
//   bucket(x) = min(Q-1, (int)((x - lo) * inv_w))        (lo / hi = smallest / largest positive of the chunk)
// is evaluated with the SAME fp32 instruction sequence for positives and for scores, and every step of it (correctly
// rounded subtract, multiply by a positive constant, truncation, clamp) is monotone non-decreasing.  Hence
// bucket(s) > bucket(p) implies s > p and bucket(s) < bucket(p) implies s < p: for a score in bucket q,
//   #{positives <= s} = #{p : bucket(p) < q} + #{p in bucket q : p <= s}
// EXACTLY -- no rounding margin is needed, and only the (at most m of Q) buckets that hold a positive cost compares.
// table[q] = #{p : bucket(p) < q} | (#{p : bucket(p) == q} << 16).
constexpr int kAucChunk = 4000;     // positives staged in shared memory per sweep (16 KB)
constexpr int kAucBuckets = 8192;   // 32 KB table: with the positives 48 KB per block, 4 blocks per SM
constexpr int kAucLinear = 8;       // positives compared one by one inside a bucket; more than that -> binary search

__device__ __forceinline__ int auc_bucket(float x, float lo, float inv_w) {
  const int q = (int)__fmul_rn(__fsub_rn(x, lo), inv_w);
  return q < kAucBuckets - 1 ? q : kAucBuckets - 1;
}

__global__ void __launch_bounds__(256) auc_rows_kernel(const float* __restrict__ scores, int64_t ld, int row0,
                                                       int64_t n_posts, const int32_t* __restrict__ labels,
                                                       const int64_t* __restrict__ seg_ptr,
                                                       const float* __restrict__ pos_sorted,
                                                       const float* __restrict__ best_score,
                                                       const int32_t* __restrict__ best_index, int64_t index_base,
                                                       unsigned long long* __restrict__ auc_num,
                                                       unsigned long long* __restrict__ before_first) {
  extern __shared__ uint32_t auc_smem[];
  uint32_t* table = auc_smem;                                             // [kAucBuckets]
  float* spos = reinterpret_cast<float*>(auc_smem + kAucBuckets);          // [kAucChunk + 1], +inf sentinel behind the last
  __shared__ unsigned long long red[2][8];
  const int r = blockIdx.x;
  const int b = row0 + r;
  const int64_t c0 = n_posts * blockIdx.y / gridDim.y, c1 = n_posts * (blockIdx.y + 1) / gridDim.y;
  const int64_t p0 = seg_ptr[b], np = seg_ptr[b + 1] - p0;
  if (np == 0) return;
  const float bs = best_score[b];
  const int bi = best_index[b];
  const float* row = scores + (int64_t)r * ld;
  unsigned long long auc = 0ull, before = 0ull;
  for (int64_t ch = 0; ch < np; ch += kAucChunk) {
    const int m = (int)((np - ch) < kAucChunk ? (np - ch) : kAucChunk);
    __syncthreads();
    for (int i = threadIdx.x; i <= m; i += blockDim.x) spos[i] = i < m ? pos_sorted[p0 + ch + i] : INFINITY;
    __syncthreads();
    const float lo_s = spos[0], hi_s = spos[m - 1];
    // hi == lo (one positive, or all tied): no score is routed through the table (see the range tests below)
    const float inv_w = hi_s > lo_s ? (float)kAucBuckets / (hi_s - lo_s) : 0.f;
    {
      // thread t owns buckets [q0, q1): the first positive whose bucket is >= q0 by binary search (bucket() is monotone
      // over the sorted positives), then one walk over its buckets
      constexpr int per = kAucBuckets / 256;
      const int q0 = threadIdx.x * per;
      int lo = 0, hi = m;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (auc_bucket(spos[mid], lo_s, inv_w) >= q0) hi = mid; else lo = mid + 1; }
      int i = lo;
      for (int q = q0; q < q0 + per; ++q) {
        const int first = i;
        while (i < m && auc_bucket(spos[i], lo_s, inv_w) == q) ++i;
        table[q] = (uint32_t)first | ((uint32_t)(i - first) << 16);
      }
    }
    __syncthreads();
    // 8 independent (score, label) loads and 8 independent table lookups in flight per thread, then a branch-free
    // resolve.  No range tests are needed: a score below every positive lands (clamped) in bucket 0, whose first
    // positive is lo > s, so it counts 0; a score at or above every positive lands in the last bucket and passes all of
    // its positives, so it counts m.  Only a NaN score needs care (`e > NaN` is False: it must add nothing).
    constexpr int U = 8;
    const float* rowp = row + c0;
    const int32_t* labp = labels + c0;
    const int len = (int)(c1 - c0);                          // < 2^31: the caller checks that global indices fit int32
    const int gbase = (int)(index_base + c0);
    const bool first_chunk = ch == 0;
    auto batch = [&](int j0, auto checked) {
      constexpr bool CHECK = decltype(checked)::value;       // last, partial batch of the thread: bounds-checked loads
      float sv[U];
      int lv[U];
      uint32_t tv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * 256;
        const bool ok = !CHECK || j < len;
        sv[u] = ok ? __ldcs(rowp + j) : __int_as_float(0x7FC00000);   // streamed once: evict first, labels stay in L2
        lv[u] = ok ? __ldg(labp + j) : b;                    // out of range: NaN score, the brand's own label -> adds nothing
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = auc_bucket(sv[u], lo_s, inv_w);        // (int)NaN = 0, below lo -> negative: clamped to bucket 0
        tv[u] = table[q > 0 ? q : 0];
      }
      unsigned int auc32 = 0, before32 = 0;                  // <= 8 * 4000 per batch: no 32-bit overflow
      uint32_t slow = 0;                                     // slots whose bucket holds SEVERAL positives (rare)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float s = sv[u];
        if (first_chunk) {
          before32 += s > bs ? 1u : 0u;
          if (s == bs) before32 += (gbase + j0 + u * 256) < bi ? 1u : 0u;      // ties with the best positive: rare
        }
        // idx = #{positives of this chunk <= s} = table count + one compare against the bucket's first positive
        // (spos[m] = +inf pads the end); a negative adds m - idx
        const int first = (int)(tv[u] & 0xFFFFu);
        const int n = (int)(tv[u] >> 16);
        const int idx = first + ((n != 0 && spos[first] <= s) ? 1 : 0);
        slow |= n > 1 ? (1u << u) : 0u;
        auc32 += (lv[u] != b && s == s) ? (unsigned int)(m - idx) : 0u;        // negatives only (evaluator.py:112)
      }
      while (slow) {                                         // redo these slots exactly: walk / search inside the bucket
        const int u = __ffs(slow) - 1;
        slow &= slow - 1;
        float s = sv[0]; int l = lv[0]; uint32_t t = tv[0];
#pragma unroll
        for (int i = 1; i < U; ++i) { s = i == u ? sv[i] : s; l = i == u ? lv[i] : l; t = i == u ? tv[i] : t; }
        const int first = (int)(t & 0xFFFFu);
        int n = (int)(t >> 16), idx = first;
        if (n <= kAucLinear) {
          while (n > 0 && spos[idx] <= s) { ++idx; --n; }
        } else {                                             // crowded bucket (ties / clustered positives): binary search
          int lo = first, hi = first + n;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (spos[mid] > s) hi = mid; else lo = mid + 1; }
          idx = lo;
        }
        const int quick = first + (spos[first] <= s ? 1 : 0);        // what the branch-free pass counted for this slot
        if (l != b && s == s) auc32 -= (unsigned int)(idx - quick);   // m - idx instead of m - quick (idx >= quick)
      }
      auc += auc32; before += before32;
    };
    int j0 = threadIdx.x;
    for (; j0 + (U - 1) * 256 < len; j0 += U * 256) batch(j0, std::false_type{});
    if (j0 < len) batch(j0, std::true_type{});
  }
  // block reduce
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    auc += __shfl_xor_sync(0xffffffffu, auc, o);
    before += __shfl_xor_sync(0xffffffffu, before, o);
  }
  if (lane == 0) { red[0][warp] = auc; red[1][warp] = before; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0, c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; c += red[1][w]; }
    if (a) atomicAdd(auc_num + b, a);
    if (c) atomicAdd(before_first + b, c);
  }
}

// One [rows, nb] int64 block for the single device->host copy of an evaluation:
//   0 n_pos   1 first rank inside the top-k list (-1 = none)   2 count before the first positive   3 row 2 is valid
//   4 hit mask (64 relevance bits)   5 AUC numerator (when given)
// Row 3: all_valid ? 1 : (first_in_list < 0 && n_pos > 0)  -- the brands whose first positive fell outside the list.
__global__ void pack_stats_kernel(const int32_t* __restrict__ n_pos, const int32_t* __restrict__ first_in_list,
                                  const unsigned long long* __restrict__ before_first,
                                  const unsigned long long* __restrict__ hit_mask,
                                  const unsigned long long* __restrict__ auc_num, int nb, int all_valid,
                                  long long* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int np_ = n_pos[b], fi = first_in_list[b];
  out[b] = np_;
  out[(size_t)nb + b] = fi;
  out[(size_t)2 * nb + b] = (long long)before_first[b];
  out[(size_t)3 * nb + b] = all_valid ? 1 : ((fi < 0 && np_ > 0) ? 1 : 0);
  out[(size_t)4 * nb + b] = (long long)hit_mask[b];
  if (auc_num) out[(size_t)5 * nb + b] = (long long)auc_num[b];
}

// thr_index[b] = best_index[b] where the first positive is missing from the list, else -1 (the count pass skips it)
__global__ void missing_threshold_kernel(const int32_t* __restrict__ n_pos, const int32_t* __restrict__ first_in_list,
                                         const int32_t* __restrict__ best_index, int nb, int32_t* __restrict__ thr_index) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  thr_index[b] = (first_in_list[b] < 0 && n_pos[b] > 0) ? best_index[b] : -1;
}

// Multi-GPU exchange, after the all-gather of every shard's (n_pos, best positive): n_pos = sum over shards, best
// positive = the largest (score, ~index) key over the shards that have one.  Shard g's arrays start `stride` 32-bit
// words after shard g-1's (they live inside the packed all-gather buffer).
__global__ void reduce_shard_stats_kernel(const int32_t* __restrict__ n_pos_g, const float* __restrict__ best_s_g,
                                          const int32_t* __restrict__ best_i_g, int g, int nb, int64_t stride,
                                          int32_t* __restrict__ n_pos, float* __restrict__ best_score,
                                          int32_t* __restrict__ best_index) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  int total = 0;
  unsigned long long best = 0ull;
  for (int r = 0; r < g; ++r) {
    total += n_pos_g[(int64_t)r * stride + b];
    const int32_t idx = best_i_g[(int64_t)r * stride + b];
    if (idx >= 0) {
      const unsigned long long key = make_key(best_s_g[(int64_t)r * stride + b], (uint32_t)idx);
      best = key > best ? key : best;
    }
  }
  n_pos[b] = total;
  if (best != 0ull) {
    best_score[b] = key_score(best);
    best_index[b] = (int32_t)key_index(best);
  } else {
    best_score[b] = -INFINITY;
    best_index[b] = -1;
  }
}

// ---- A10: util/metric.py scorers (P@k, AP@k, RR, NDCG@k, DCG@k) over batches of sorted label lists ------------
// One warp per list.  Integer label lists (graded relevance) in, float64 scores out, bit-identical to the reference's
// Python arithmetic: every float64 operation is done in the reference's order by one lane, the warp's job is to find the
// positions that matter (ballots over 32 labels at a time: zero-gain positions add +0.0 and are skipped) and the
// descending order of the grades for the ideal DCG.  `log2_table[i]` = math.log(i, 2) computed on the HOST (CPython's
// two-argument log is log(x) / log(2), which no device intrinsic reproduces bit for bit).
enum { SCORER_P = 0, SCORER_AP = 1, SCORER_RR = 2, SCORER_NDCG = 3, SCORER_DCG = 4 };

__device__ __forceinline__ int warp_max_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}

// util/metric.py:80-88 (NDCGScorer.getDCG) over the first `length` labels, in list order
__device__ double ndcg_dcg_in_order(const int32_t* lab, int length, const double* log2_table, int lane) {
  double dcg = 0.0;
  for (int base = 0; base < length; base += 32) {
    const int i = base + lane;
    const int g = i < length ? max(lab[i], 0) : 0;
    uint32_t m = __ballot_sync(0xffffffffu, g > 0);
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const int gl = __shfl_sync(0xffffffffu, g, l);
      const int pos = base + l;
      if (pos == 0) dcg = (double)gl;                                   // dcg = max(sorted_labels[0], 0)
      else dcg += (double)gl / log2_table[pos + 1];                     // float(rel) / math.log(i + 1, 2)
    }
  }
  return dcg;
}

__global__ void __launch_bounds__(256) scorer_kernel(const int32_t* __restrict__ labels, int64_t ld,
                                                     const int32_t* __restrict__ lengths, int n_lists, int max_len, int kind,
                                                     int k, const double* __restrict__ log2_table, double* __restrict__ out) {
  const int list = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (list >= n_lists) return;
  const int32_t* lab = labels + (int64_t)list * ld;
  const int n = lengths ? min(max(lengths[list], 0), max_len) : max_len;
  const int length = (k > 0 && k <= n) ? k : n;                         // MetricScorer.getLength
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  double res = 0.0;
  if (kind == SCORER_P) {                                               // util/metric.py:58-68
    int rel = 0;
    for (int i = lane; i < length; i += 32) rel += lab[i] >= 1 ? 1 : 0;
    rel = (int)__reduce_add_sync(0xffffffffu, (unsigned)rel);
    res = length > 0 ? (double)rel / (double)length : nan;
  } else if (kind == SCORER_RR) {                                       // util/metric.py:49-55 (whole list, not @k)
    res = 0.0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const uint32_t m = __ballot_sync(0xffffffffu, i < n && lab[i] >= 1);
      if (m) { res = 1.0 / (double)(base + __ffs(m)); break; }
    }
  } else if (kind == SCORER_AP) {                                       // util/metric.py:26-45
    int nr = 0;
    for (int i = lane; i < n; i += 32) nr += lab[i] > 0 ? 1 : 0;
    nr = (int)__reduce_add_sync(0xffffffffu, (unsigned)nr);
    double ap = 0.0;
    int rel = 0;
    for (int base = 0; base < length && nr > 0; base += 32) {
      const int i = base + lane;
      uint32_t m = __ballot_sync(0xffffffffu, i < length && lab[i] >= 1);
      while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        rel += 1;
        ap += (double)rel / ((double)(base + l) + 1.0);
      }
    }
    res = nr > 0 ? ap / (double)nr : 0.0;
  } else if (kind == SCORER_NDCG) {                                     // util/metric.py:71-92
    if (n == 0) res = nan;                                              // sorted_labels[0] raises IndexError there
    else {
      const double d = ndcg_dcg_in_order(lab, length, log2_table, lane);
      // ideal: the same sum over sorted(labels, reverse=True)[:length]; only positive grades add anything, so walk the
      // distinct positive grades in descending order (a handful) and emit each as many times as it occurs
      double ideal = 0.0;
      int pos = 0, bound = 0x7FFFFFFF;
      while (pos < length) {
        int g = 0, c = 0;
        for (int i = lane; i < n; i += 32) { const int v = lab[i]; if (v < bound && v > g) g = v; }
        g = warp_max_i32(g);
        if (g <= 0) break;
        for (int i = lane; i < n; i += 32) c += lab[i] == g ? 1 : 0;
        c = (int)__reduce_add_sync(0xffffffffu, (unsigned)c);
        for (int t = 0; t < c && pos < length; ++t, ++pos) {
          if (pos == 0) ideal = (double)g; else ideal += (double)g / log2_table[pos + 1];
        }
        bound = g;
      }
      res = d / ideal;                                                  // 0 / 0 -> NaN: the reference raises ZeroDivisionError
    }
  } else {                                                              // SCORER_DCG, util/metric.py:95-116
    // 0.01757 * sum(parts), parts over sorted_labels[:k]; CPython >= 3.12 sums floats with Neumaier compensation
    const int cnt = k < n ? (k > 0 ? k : 0) : n;
    double total = 0.0, comp = 0.0;
    for (int i = 0; i < cnt; ++i) {
      const double x = (ldexp(1.0, lab[i]) - 1.0) / log2_table[i + 2];  // (2**rel - 1) / math.log(index + 1, 2), index = i + 1
      const double t = total + x;
      if (fabs(total) >= fabs(x)) comp += (total - t) + x; else comp += (x - t) + total;
      total = t;
    }
    if (comp != 0.0 && isfinite(comp)) total += comp;
    res = 0.01757 * total;
  }
  if (lane == 0) out[list] = res;
}

}  // namespace frx

extern "C" {

int frx_metric_scores(const int32_t* labels, int64_t ld, const int32_t* lengths, int n_lists, int max_len, int kind, int k,
                      const double* log2_table, double* out, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(labels && log2_table && out, "frx_metric_scores: NULL pointer");
  FRX_CHECK_ARG(n_lists > 0 && max_len >= 0 && ld >= max_len, "frx_metric_scores: bad sizes");
  FRX_CHECK_ARG(kind >= SCORER_P && kind <= SCORER_DCG && k >= 0, "frx_metric_scores: unknown scorer %d / k %d", kind, k);
  int dev = 0;
  FRX_CUDA(cudaGetDevice(&dev));
  const int rc = frx_device_check(dev);
  if (rc) return rc;
  const int warps_per_block = 8;
  scorer_kernel<<<(n_lists + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, (cudaStream_t)stream>>>(
      labels, ld, lengths, n_lists, max_len, kind, k, log2_table, out);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}


int frx_reduce_shard_stats(const int32_t* n_pos_g, const float* best_score_g, const int32_t* best_index_g, int g, int nb,
                           int64_t stride_words, int32_t* n_pos, float* best_score, int32_t* best_index, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(n_pos_g && best_score_g && best_index_g && n_pos && best_score && best_index, "frx_reduce_shard_stats: NULL pointer");
  FRX_CHECK_ARG(g >= 1 && nb >= 1 && stride_words >= nb, "frx_reduce_shard_stats: bad sizes");
  reduce_shard_stats_kernel<<<(nb + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_pos_g, best_score_g, best_index_g, g, nb,
                                                                                 stride_words, n_pos, best_score, best_index);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_pack_rank_stats(const int32_t* n_pos, const int32_t* first_in_list, const unsigned long long* before_first,
                        const unsigned long long* hit_mask, const unsigned long long* auc_num, int nb, int all_valid,
                        long long* out, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(n_pos && first_in_list && before_first && hit_mask && out, "frx_pack_rank_stats: NULL pointer");
  FRX_CHECK_ARG(nb > 0, "frx_pack_rank_stats: bad size");
  pack_stats_kernel<<<(nb + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_pos, first_in_list, before_first, hit_mask, auc_num,
                                                                         nb, all_valid, out);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_missing_thresholds(const int32_t* n_pos, const int32_t* first_in_list, const int32_t* best_index, int nb,
                           int32_t* thr_index, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(n_pos && first_in_list && best_index && thr_index, "frx_missing_thresholds: NULL pointer");
  FRX_CHECK_ARG(nb > 0, "frx_missing_thresholds: bad size");
  missing_threshold_kernel<<<(nb + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_pos, first_in_list, best_index, nb, thr_index);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_label_stats(const int32_t* labels, const float* pos_score, int64_t n_posts, int nb, int64_t index_base,
                    int32_t* n_pos, float* best_score, int32_t* best_index, void* workspace_nb_u64, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(labels && pos_score && n_pos && best_score && best_index && workspace_nb_u64, "frx_label_stats: NULL pointer");
  FRX_CHECK_ARG(nb > 0 && n_posts >= 0, "frx_label_stats: bad sizes");
  FRX_CHECK_ARG(index_base >= 0 && index_base + n_posts <= 2147483647LL, "frx_label_stats: index range must fit int32");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(workspace_nb_u64);
  FRX_CUDA(cudaMemsetAsync(n_pos, 0, (size_t)nb * sizeof(int32_t), st));
  FRX_CUDA(cudaMemsetAsync(keys, 0, (size_t)nb * sizeof(unsigned long long), st));
  if (n_posts > 0) {
    int64_t blocks = (n_posts + 255) / 256;
    const int64_t maxb = (int64_t)num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    if (nb <= kPrivBrands && n_posts >= 65536) {
      int64_t pb = (n_posts + 4095) / 4096;                        // >= 4 k posts per block
      if (pb > num_sms()) pb = num_sms();
      label_stats_priv_kernel<<<(int)pb, 1024, 0, st>>>(labels, pos_score, n_posts, nb, index_base, n_pos, keys);
    } else {
      label_stats_kernel<<<(int)blocks, 256, 0, st>>>(labels, pos_score, n_posts, nb, index_base, n_pos, keys);
    }
    FRX_LAUNCH_CHECK();
  }
  decode_best_kernel<<<(nb + 255) / 256, 256, 0, st>>>(keys, n_pos, nb, best_score, best_index);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_rank_from_topk(const int32_t* topk_index, int nb, int k, const int32_t* labels, int64_t n_posts,
                       int64_t index_base, unsigned long long* hit_mask, int32_t* first_rank, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(topk_index && labels && hit_mask && first_rank, "frx_rank_from_topk: NULL pointer");
  FRX_CHECK_ARG(nb > 0 && k > 0, "frx_rank_from_topk: bad sizes");
  const int warps_per_block = 8;
  rank_from_topk_kernel<<<(nb + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, (cudaStream_t)stream>>>(
      topk_index, nb, k, labels, n_posts, index_base, hit_mask, first_rank);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_group_positives(const int32_t* labels, const float* pos_score, int64_t n_posts, int nb, const int32_t* n_pos,
                        int64_t* seg_ptr, float* pos_sorted, void* workspace_nb_i64, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(labels && pos_score && n_pos && seg_ptr && pos_sorted && workspace_nb_i64, "frx_group_positives: NULL pointer");
  FRX_CHECK_ARG(nb > 0 && n_posts >= 0, "frx_group_positives: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* cursor = reinterpret_cast<unsigned long long*>(workspace_nb_i64);
  seg_scan_kernel<<<1, 1024, 0, st>>>(n_pos, nb, seg_ptr, cursor);
  FRX_LAUNCH_CHECK();
  if (n_posts > 0) {
    int64_t blocks = (n_posts + 255) / 256;
    const int64_t maxb = (int64_t)num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    scatter_pos_kernel<<<(int)blocks, 256, 0, st>>>(labels, pos_score, n_posts, nb, seg_ptr, cursor, pos_sorted);
    FRX_LAUNCH_CHECK();
    const size_t smem = (size_t)kSortSmemFloats * sizeof(float);
    FRX_CUDA(cudaFuncSetAttribute(sort_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sort_segments_kernel<<<nb, 256, smem, st>>>(seg_ptr, pos_sorted);
    FRX_LAUNCH_CHECK();
  }
  return FRX_OK;
}

int frx_auc_rows(const float* scores, int64_t ld, int row0, int n_rows, int64_t n_posts, const int32_t* labels,
                 const int64_t* seg_ptr, const float* pos_sorted, const float* best_score, const int32_t* best_index,
                 int64_t index_base, unsigned long long* auc_num, unsigned long long* before_first, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(scores && labels && seg_ptr && pos_sorted && best_score && best_index && auc_num && before_first,
                "frx_auc_rows: NULL pointer");
  FRX_CHECK_ARG(n_rows > 0 && n_posts > 0 && ld >= n_posts && row0 >= 0, "frx_auc_rows: bad sizes");
  // ~8 blocks per resident slot (4 per SM): short enough for the block scheduler to even out the tail, long enough
  // (>= 32 k scores) that rebuilding the row's bucket table per block stays a few percent
  int64_t ysplit = ((int64_t)num_sms() * 32 + n_rows - 1) / n_rows;
  const int64_t max_split = (n_posts + 32767) / 32768;
  if (ysplit > max_split) ysplit = max_split;
  if (ysplit < 1) ysplit = 1;
  if (ysplit > 65535) ysplit = 65535;
  dim3 grid(n_rows, (unsigned)ysplit);
  const size_t smem = (size_t)kAucBuckets * sizeof(uint32_t) + (size_t)(kAucChunk + 1) * sizeof(float);
  FRX_CUDA(cudaFuncSetAttribute(auc_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  auc_rows_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(scores, ld, row0, n_posts, labels, seg_ptr, pos_sorted,
                                                          best_score, best_index, index_base, auc_num, before_first);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

}  // extern "C"
