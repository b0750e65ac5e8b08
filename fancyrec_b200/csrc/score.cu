// A5 + A6: brand x post cosine-score contraction on the 5th-gen tensor cores with the per-brand
// top-k selection fused into the GEMM epilogue (the score matrix is never written).
//
// Replaces cal_sim's im.mm(s.t()) (evaluator.py:29), the device->host copy of the full score matrix
// (evaluator.py:96) and the per-brand sorted()/np.argsort (evaluator.py:108-109,124).
//
// Shape of the kernel (one persistent CTA per SM, warp-specialised):
//   warp 0 (one lane)  TMA producer: A = 128 brands x 64 k, B = 256 posts x 64 k (bf16, SWIZZLE_128B)
//                      into a 4-stage shared-memory ring (48 KB / stage)
//   warp 1 (one lane)  tcgen05.mma issuer: M=128, N=256, K=16 per instruction, fp32 accumulators in
//                      TMEM, two accumulator stages (2 x 256 columns = all 512 TMEM columns) so the
//                      epilogue of tile t overlaps the main loop of tile t+1
//   warp 2             TMEM alloc / dealloc
//   warps 4-7          epilogue: thread = TMEM lane = one brand row.  Modes:
//       TOPK   per-row running threshold in a register; a score >= threshold is appended (rare after
//              warm-up) as a packed 64-bit (score, ~index) key to the row's private candidate buffer;
//              a full buffer is compacted warp-cooperatively by an MSB-first radix select that raises
//              the threshold.  Also extracts S[label[j], j] (the positive's score) for the metric side.
//       DENSE  writes the fp32 tile (score-tolerance tests, AUC row sweep).
//       COUNT  counts, per row, the scores that precede a given (score, index) threshold
//              (= rank of the first positive, evaluator.py:116) without materialising anything.
// A work item is (m_tile, split): 128 brand rows x a contiguous range of 256-post tiles.  Items are
// numbered split-major so that the CTAs resident at any time read the same post range (served by L2).
// Per-item candidate lists are merged by merge_partials_kernel (bitonic sort of <= 16384 keys in smem).
#include <stdlib.h>
#include "common.cuh"
#include "sm100.cuh"

namespace frx {
using namespace sm100;

constexpr int BM = 128, BN = 256;
// one k-block = one 128-byte swizzle row per operand row: 64 bf16 or 32 tf32 (fp32 storage); 4 MMAs per k-block
// (K = 16 bf16 / 8 tf32 = 32 bytes each), so stage bytes and the per-MMA descriptor advance are the same in both modes
constexpr int BK_BYTES = 128, MMAS_PER_KBLOCK = 4;
constexpr int A_STAGE_BYTES = BM * BK_BYTES;   // 16 KB
constexpr int MAX_STAGES = 6;
// Shared-memory ring.  Single CTA: 4 stages of {A 128 x 64, B 256 x 64} = 48 KB.  CTA pair (cta_group::2, one 256 x 256
// MMA tile per pair): each CTA stages its own 128 rows of A and HALF of the B tile, 32 KB per stage -> 6 stages in the
// same 192 KB, and a third less L2 -> shared-memory operand traffic per flop.
template <bool PAIR>
struct Ring {
  static constexpr int STAGES = PAIR ? 6 : 4;
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;
  static constexpr int B_BYTES = B_ROWS * BK_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_BYTES;
};
constexpr int RING_BYTES = Ring<false>::STAGES * Ring<false>::STAGE_BYTES;   // 192 KB in both layouts
static_assert(RING_BYTES == Ring<true>::STAGES * Ring<true>::STAGE_BYTES, "both ring layouts use the same shared memory");
constexpr int EPI_WARP0 = 4;                 // warps 4..11: lane quarter = warp % 4, column half = (warp-4)/4
// Epilogue warps: 4 (one per TMEM lane quarter, each draining all 256 columns of a tile; the default "slim" CTA of 256
// threads / 30 720 registers) or 8 (two per quarter, 128 columns each; the "wide" CTA of 384 threads / 52 224 registers).
// With 4 warps the epilogue of a tile is 8-11 k of its 28 k cycles at k <= 256 and still hides behind the next tile's MMAs
// (config 2: step 8.0 ms against 8.15-8.35 ms with 8 warps, and two finalise blocks fit next to the CTA);
// at k = 1000 the appends and compactions of the top-k epilogue need the 8 warps (config 4 on one GPU: 1.03 s against
// 1.15 s), so the fused top-k launch picks the wide kernel for k > 256.
constexpr int NUM_EPI_WARPS = 4, WIDE_EPI_WARPS = 8;
constexpr int NUM_THREADS = (EPI_WARP0 + NUM_EPI_WARPS) * 32;    // 256
constexpr int WIDE_THREADS = (EPI_WARP0 + WIDE_EPI_WARPS) * 32;  // 384
constexpr int WIDE_K = 256;                                       // fused top-k with k above this runs on the wide kernel
constexpr int TMEM_COLS = 512;
// Register budget.  A CTA is launched with a uniform count per thread, then the warp group of the TMA / MMA / TMEM
// warps hands its surplus to the epilogue warp group(s) (setmaxnreg): 128 x 72 + 128 x 168 = 256 x 120 (slim),
// 128 x 72 + 256 x 168 = 384 x 136 (wide).
constexpr int LEAN_REGS = 72, EPI_REGS = 168;
constexpr int KERNEL_REGS = (128 * LEAN_REGS + NUM_EPI_WARPS * 32 * EPI_REGS) / NUM_THREADS;        // 120
constexpr int WIDE_REGS = (128 * LEAN_REGS + WIDE_EPI_WARPS * 32 * EPI_REGS) / WIDE_THREADS;        // 136
static_assert(128 * LEAN_REGS + WIDE_EPI_WARPS * 32 * EPI_REGS == WIDE_THREADS * WIDE_REGS, "register budget is redistributed exactly");
static_assert(128 * LEAN_REGS + NUM_EPI_WARPS * 32 * EPI_REGS == NUM_THREADS * KERNEL_REGS, "register budget is redistributed exactly");
constexpr int MAX_MERGE_KEYS = 16384;
constexpr int MAX_NEED_TILES = 1024;          // COUNT mode tracks per-m-tile skip flags for up to 131072 brands
constexpr int64_t kSampleMinPosts = 262144;   // below this the warm-up is too short to be worth a sample pass
// Global per-row candidate histogram (TOPK): bin = (ordered(score) - ordered(sample threshold)) >> HIST_SHIFT, i.e.
// 64 bins over two octaves of the score above the seeded threshold (32 per octave: the refined threshold sits at most
// ~3 % below the true k-th best); the last bin also holds everything beyond.
constexpr int HIST_BINS = 64, HIST_SHIFT = 18;
constexpr int REFINE_EVERY = 8;               // tiles between two threshold refinements of a (row, column half)

enum Mode { MODE_TOPK = 0, MODE_DENSE = 1, MODE_COUNT = 2 };

struct ScoreParams {
  int nb;
  int64_t n_posts;
  int num_k_blocks, num_m_tiles, splits;
  int k_splits;                    // DENSE only: K range split over work items, partial tiles reduced afterwards
  int64_t partial_stride;          // elements between the partial tiles of consecutive k-splits
  int64_t num_n_tiles;
  int64_t index_base;
  // TOPK
  int k, cap, keep_limit;
  unsigned long long* part_keys;   // [items][column ranges: 1 slim / 2 wide][128][cap]
  int* part_cnt;                   // [items][column ranges][128]
  uint32_t* row_thr;               // [nb] best published lower bound of each row's k-th best score (ordered)
  uint32_t* row_hist;              // [nb][HIST_BINS] scores appended so far by ANY CTA, binned above row_base (nullptr = off)
  const uint32_t* row_base;        // [nb] ordered threshold seeded by the sample pass = origin of the bins (0 = row off)
  int refine_every;                // tiles between two threshold refinements of a (row, column half)
  const int32_t* labels;
  float* pos_score;
  // DENSE
  float* dense;
  int64_t ld_dense;
  float dense_scale;               // accumulators are multiplied by this before being written (1 = as is)
  // COUNT
  const float* thr_score;
  const int32_t* thr_index;
  unsigned long long* count_out;
  long long* trace;                // FRX_TRACE builds only: per-tile clock64 stamps of CTA 0 (4 per tile)
  long long* cta_trace;            // FRX_TRACE builds only: per CTA {smid, globaltimer at start, at tile 50, at the end}
};

struct SmemTail {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  // per epilogue warp: 256-bin radix histogram of the select (first 1 KB), or the 32 x 16 fp32 transpose stage of the
  // dense store (all 2 KB) -- never live at the same time
  alignas(16) uint32_t scratch[WIDE_EPI_WARPS][512];
  uint8_t tile_need[MAX_NEED_TILES];   // COUNT: m-tile has at least one row with a threshold (others are skipped)
};
constexpr size_t SMEM_BYTES = 1024 /* alignment slack */ + (size_t)RING_BYTES + sizeof(SmemTail);

// ---------------------------------------------------------------------------------------------
// Warp-cooperative MSB-first radix select over one row's candidate keys.
// Keeps the best `k` keys (or, when !exact, between k and keep_limit keys as soon as a digit boundary
// allows it), compacts them to buf[0..kept) and returns kept; *thr_out = a score that every kept key
// reaches and no dropped key exceeds (the new append threshold).
// KPL > 0: the row's keys (n <= 32*KPL) are loaded ONCE into registers (KPL independent 8-byte loads in
// flight per lane) and every pass runs from registers.  KPL == 0: generic path for large buffers, keys
// re-read from L2 each pass in batches of 8 independent loads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void select_digit(uint32_t* hist, int lane, int need, int& digit, int& above, int& bucket) {
  uint32_t h[8], lsum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { h[i] = hist[lane * 8 + i]; lsum += h[i]; }
  uint32_t suf = lsum;                         // inclusive suffix sum over lanes (higher lane = higher digits)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += t;
  }
  const bool cross = suf >= (uint32_t)need && (suf - lsum) < (uint32_t)need;
  const int cl = __ffs(__ballot_sync(0xffffffffu, cross)) - 1;   // exactly one lane crosses
  int dg = 0, ab = 0, bc = 0;
  if (lane == cl) {
    uint32_t acc = suf - lsum;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      if (bc == 0) {
        if (acc + h[i] >= (uint32_t)need) { dg = lane * 8 + i; ab = (int)acc; bc = (int)h[i]; }
        else acc += h[i];
      }
    }
  }
  digit = __shfl_sync(0xffffffffu, dg, cl);
  above = __shfl_sync(0xffffffffu, ab, cl);
  bucket = __shfl_sync(0xffffffffu, bc, cl);
}

template <int KPL>
__device__ __forceinline__ int warp_select(unsigned long long* buf, int n, int k, int keep_limit, bool exact,
                                           uint32_t* hist, float* thr_out) {
  const int lane = threadIdx.x & 31;
  unsigned long long keys[KPL > 0 ? KPL : 1];
  if (KPL > 0) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const int i = j * 32 + lane;
      keys[j] = i < n ? __ldcg(buf + i) : 0ull;    // 0 is below every real key and never matches a prefix > 0
    }
  }
  unsigned long long prefix = 0;
  int need = k, above_total = 0, shift = 56, bucket = 0;
  for (int pass = 0; pass < 8; ++pass, shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0;
    __syncwarp();
    if (KPL > 0) {
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const unsigned long long key = keys[j];
        if (j * 32 + lane < n && (pass == 0 || (key >> (shift + 8)) == prefix))
          atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
      }
    } else {
      for (int base = 0; base < n; base += 512) {          // 16 independent 8-byte loads in flight per lane
        unsigned long long kk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { const int i = base + j * 32 + lane; kk[j] = i < n ? __ldcg(buf + i) : 0ull; }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (base + j * 32 + lane < n && (pass == 0 || (kk[j] >> (shift + 8)) == prefix))
            atomicAdd(&hist[(uint32_t)(kk[j] >> shift) & 255u], 1u);
        }
      }
    }
    __syncwarp();
    int digit, above;
    select_digit(hist, lane, need, digit, above, bucket);
    above_total += above;
    need -= above;
    prefix = (prefix << 8) | (unsigned long long)digit;
    if ((!exact && above_total + bucket <= keep_limit) || pass == 7) break;
  }
  const unsigned long long thr_key = prefix << shift;   // smallest key of the boundary bucket
  int out = 0;
  if (KPL > 0) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      if (j * 32 < n) {                                  // warp-uniform
        const bool keep = (j * 32 + lane < n) && keys[j] >= thr_key;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) buf[out + __popc(m & ((1u << lane) - 1u))] = keys[j];
        out += __popc(m);
      }
    }
  } else {
    for (int base = 0; base < n; base += 512) {
      unsigned long long kk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { const int i = base + j * 32 + lane; kk[j] = i < n ? __ldcg(buf + i) : 0ull; }
      __syncwarp();                                        // every lane holds its 16 keys before anything is overwritten
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool keep = (base + j * 32 + lane < n) && kk[j] >= thr_key;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) buf[out + __popc(m & ((1u << lane) - 1u))] = kk[j];   // out + rank <= source index: unread keys are safe
        out += __popc(m);
      }
      __syncwarp();
    }
  }
  float t = ordered_to_score((uint32_t)(thr_key >> 32));
  if (t != t) t = -INFINITY;   // bucket edge decoded to a NaN pattern
  *thr_out = t;
  return out;
}

__device__ __forceinline__ int select_dispatch(int cap, unsigned long long* buf, int n, int k, int keep_limit,
                                               bool exact, uint32_t* hist, float* thr_out) {
  if (cap <= 1024) return warp_select<32>(buf, n, k, keep_limit, exact, hist, thr_out);
  return warp_select<0>(buf, n, k, keep_limit, exact, hist, thr_out);
}

// Threshold refinement from the row's global histogram: the highest bin whose suffix count reaches k.  At least k
// posts seen so far (by any CTA) score at or above that bin's lower edge, so the edge is a valid lower bound of the
// row's global k-th best score -- every CTA of the row may drop everything below it.  -1 = fewer than k counted.
__device__ __forceinline__ int hist_edge(const uint32_t* hrow, uint32_t k) {
  uint32_t acc = 0;
  int reached = 0;                               // number of bins b with suffix(b) >= k (suffix is non-increasing in b)
#pragma unroll
  for (int half = 1; half >= 0; --half) {
    uint4 c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = __ldcg(reinterpret_cast<const uint4*>(hrow) + half * 8 + i);
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      acc += c[i].w; reached += acc >= k;
      acc += c[i].z; reached += acc >= k;
      acc += c[i].y; reached += acc >= k;
      acc += c[i].x; reached += acc >= k;
    }
    if (reached) return half * 32 + reached - 1;
  }
  return -1;
}

// One 32-column chunk of one accumulator row -> the fp32 score matrix (DENSE mode; TOPK mode when the caller also wants
// the scores, e.g. for the exact-AUC sweep).  STREAM: evict-first stores, so that a 4 GB score matrix written next to
// the fused top-k pass does not push the shared post tiles out of L2.
template <bool STREAM>
__device__ __forceinline__ void store_dense_chunk(float* dst, const uint32_t (&v)[32], int nvalid, float sc) {
  if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 o = make_float4(__uint_as_float(v[i]) * sc, __uint_as_float(v[i + 1]) * sc,
                                   __uint_as_float(v[i + 2]) * sc, __uint_as_float(v[i + 3]) * sc);
      if (STREAM) __stcs(reinterpret_cast<float4*>(dst + i), o);
      else *reinterpret_cast<float4*>(dst + i) = o;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) dst[i] = __uint_as_float(v[i]) * sc;
  }
}

// The same chunk for the warp's 32 rows at once, transposed through shared memory so that every store instruction writes
// 8 rows x 64 contiguous bytes (two whole sectors per row) instead of 32 rows x 16 bytes: a quarter of the LSU wavefronts
// and no partially written sectors.  Lane r enters with its row's 32 columns in v; `stage` = 512 floats of this warp.
// Layout of one half chunk (16 columns): row r at stage[16 r ..], its four float4 groups XOR-swizzled by (r >> 1) & 3 --
// conflict-free both for the row-wise writes and for the reads (8 lanes cover rows 2j, 2j+1 completely).
// Requires: all 32 lanes, 32 valid columns, 16-byte aligned row segments (base and ld).
template <bool STREAM>
__device__ __forceinline__ void store_dense_chunk_staged(float* base, int64_t ld, const uint32_t (&v)[32], float sc,
                                                         float* stage, int lane, int rows_valid) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int phys = g ^ ((lane >> 1) & 3);
      *reinterpret_cast<float4*>(stage + lane * 16 + 4 * phys) =
          make_float4(__uint_as_float(v[half * 16 + 4 * g]) * sc, __uint_as_float(v[half * 16 + 4 * g + 1]) * sc,
                      __uint_as_float(v[half * 16 + 4 * g + 2]) * sc, __uint_as_float(v[half * 16 + 4 * g + 3]) * sc);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 8 * i + (lane >> 2), g = lane & 3;
      const float4 o = *reinterpret_cast<const float4*>(stage + r * 16 + 4 * (g ^ ((r >> 1) & 3)));
      if (r < rows_valid) {
        float* dst = base + (int64_t)r * ld + half * 16 + 4 * g;
        if (STREAM) __stcs(reinterpret_cast<float4*>(dst), o);
        else *reinterpret_cast<float4*>(dst) = o;
      }
    }
    __syncwarp();
  }
}

// CL > 1 (with PAIR = false): a cluster of CL CTAs works on CL consecutive m-tiles of the SAME post range.  Every CTA runs
// its own cta_group::1 MMAs on its own A tile, but the B tile is fetched ONCE per cluster: CTA c loads rows
// [c * 256 / CL, (c + 1) * 256 / CL) of it and TMA-multicasts them into the shared memory of all CL CTAs.  A ring slot is
// free when ALL CTAs have consumed it (multicast MMA commits, `empty` counts CL arrivals).  Sharing the post operand no
// longer depends on the CTAs of a range running in lock step so that their L2 reads coalesce.
template <int MODE, bool TF32, bool PAIR, int CL, int EW>
__device__ __forceinline__ void score_body(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const ScoreParams& P) {
  static_assert(!(PAIR && CL > 1), "the CTA-pair variant and the multicast clusters are separate variants");
  static_assert(EW == 4 || EW == 8, "one or two epilogue warps per TMEM lane quarter");
  constexpr int HV = EW / 4;                               // column ranges a tile is split into between the epilogue warps
  constexpr int NT = (EPI_WARP0 + EW) * 32;                // threads of this CTA
  constexpr bool CLUSTERED = PAIR || CL > 1;
  constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1u);
  using R = Ring<PAIR>;
  constexpr int STAGES = R::STAGES;
  constexpr int B_STAGE = R::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;            // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  SmemTail* tail = reinterpret_cast<SmemTail*>(smem + (size_t)RING_BYTES);
  const uint32_t smem_a = base, smem_b = base + STAGES * A_STAGE_BYTES;
  constexpr int BK = TF32 ? 32 : 64;                      // operand elements per k-block
  {
    // ---- prologue (every warp, launch-time register budget) ----
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
      prefetch_tmap(&tmap_a);
      prefetch_tmap(&tmap_b);
      for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&tail->full[s]), 1); mbar_init(smem_u32(&tail->empty[s]), CL); }
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&tail->tmem_full[s]), 1);
        mbar_init(smem_u32(&tail->tmem_empty[s]), EW * (PAIR ? 2 : 1));   // the leader hears both CTAs' epilogues
      }
      fence_barrier_init();
    }
    if (warp == 2) {
      if (PAIR) tmem_alloc_pair<TMEM_COLS>(smem_u32(&tail->tmem_base));
      else tmem_alloc<TMEM_COLS>(smem_u32(&tail->tmem_base));
    }
    if (MODE == MODE_COUNT) {
      // rows without a threshold (thr_index < 0) are not counted; an m-tile made only of such rows is skipped by all
      // three roles, so the pass costs only the m-tiles that need it (nothing at all when no first positive is missing)
      const int nt = P.num_m_tiles < MAX_NEED_TILES ? P.num_m_tiles : MAX_NEED_TILES;
      for (int i = threadIdx.x; i < nt; i += NT) tail->tile_need[i] = 0;
      __syncthreads();
      for (int r = threadIdx.x; r < P.nb && r < MAX_NEED_TILES * BM; r += NT)
        if (__ldg(P.thr_index + r) >= 0) tail->tile_need[r / BM] = 1;
    }
    tc_fence_before();
    if (CLUSTERED) cluster_sync_all(); else __syncthreads();    // the peers' barriers are initialised before anything targets them
    tc_fence_after();
  }
  const uint32_t tid = threadIdx.x, bid = blockIdx.x, nbid = gridDim.x;
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31);
  // CTA pair: cluster rank 0 = leader (issues the MMAs); the pair shares one scheduling slot and one 256-row m-unit
  constexpr int UNIT = PAIR ? 2 : CL;                      // m-tiles per scheduling unit (CTA, pair or cluster)
  const uint32_t crank = CLUSTERED ? cluster_ctarank() : 0u;
  const int slot = (int)bid / UNIT;
  const int nslots = (int)nbid / UNIT;
  const int m_units = P.num_m_tiles / UNIT;                // the host pads num_m_tiles to a multiple of UNIT
  const uint32_t tmem_base = tail->tmem_base;

  // item = (split * m_units + m_unit) * k_splits + ks ; a unit is one 128-row m-tile, or the pair's two m-tiles
  const int n_items = m_units * P.splits * P.k_splits;
  auto unit_skipped = [&](int mu) -> bool {                 // COUNT only; both CTAs of a pair take the same decision
    if (MODE != MODE_COUNT) return false;
    const int t0_ = mu * UNIT, t1_ = mu * UNIT + UNIT - 1;
    if (t1_ >= MAX_NEED_TILES) return false;
    bool need = false;
#pragma unroll
    for (int t_ = 0; t_ < UNIT; ++t_) need = need || tail->tile_need[t0_ + t_];
    return !need;
  };

  // Register redistribution: each role's code must be DOMINATED by its setmaxnreg -- ptxas budgets a region by the
  // setmaxnreg that dominates it (after a merge of both it assumes the smaller budget).
  if (warp == 0) {
    // =========================== TMA producer ===========================
    setmaxnreg_dec<LEAN_REGS>();
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = slot; item < n_items; item += nslots) {
        const int ks = item % P.k_splits, mi = item / P.k_splits;
        const int mu = mi % m_units, split = mi / m_units;
        if (unit_skipped(mu)) continue;
        const int m_tile = mu * UNIT + (int)crank;
        const int kb0 = P.num_k_blocks * ks / P.k_splits, kb1 = P.num_k_blocks * (ks + 1) / P.k_splits;
        const int64_t t0 = P.num_n_tiles * split / P.splits, t1 = P.num_n_tiles * (split + 1) / P.splits;
        for (int64_t t = t0; t < t1; ++t) {
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(smem_u32(&tail->empty[stage]), phase ^ 1);
            const uint32_t fb = smem_u32(&tail->full[stage]);
            if (PAIR) {
              // both CTAs' bytes are counted on the LEADER's barrier (the only thread that waits for operands is its
              // MMA issuer); this CTA brings its 128 rows of A and its half of the B tile
              if (crank == 0) mbar_arrive_expect_tx(fb, 2 * R::STAGE_BYTES);
              const uint32_t lfb = mapa_shared(fb, 0);
              tma_load_2d_pair(smem_a + stage * A_STAGE_BYTES, &tmap_a, lfb, kb * BK, m_tile * BM);
              tma_load_2d_pair(smem_b + stage * B_STAGE, &tmap_b, lfb, kb * BK, (int32_t)(t * BN + crank * (BN / 2)));
            } else if (CL > 1) {
              // this CTA's own A tile + the whole B tile, which arrives as CL multicast slices (one from every CTA)
              mbar_arrive_expect_tx(fb, R::STAGE_BYTES);
              tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_a, fb, kb * BK, m_tile * BM);
              tma_load_2d_multicast(smem_b + stage * B_STAGE + crank * (B_STAGE / CL), &tmap_b, fb, kb * BK,
                                    (int32_t)(t * BN + crank * (BN / CL)), CL_MASK);
            } else {
              mbar_arrive_expect_tx(fb, R::STAGE_BYTES);
              tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_a, fb, kb * BK, m_tile * BM);
              tma_load_2d(smem_b + stage * B_STAGE, &tmap_b, fb, kb * BK, (int32_t)(t * BN));
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (CLUSTERED) {
        // drain: every multicast "slot free" arrival aimed at this CTA has landed before it may exit
        for (int i = 0; i < STAGES; ++i) {
          mbar_wait(smem_u32(&tail->empty[stage]), phase ^ 1);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    setmaxnreg_dec<LEAN_REGS>();
    if (lane == 0 && (!PAIR || crank == 0)) {              // of a pair only the leader issues (its MMAs span both SMs);
                                                           // in a multicast cluster every CTA issues its own
      constexpr int MMA_M = PAIR ? 2 * BM : BM;
      constexpr uint32_t idesc = TF32 ? make_idesc_tf32(MMA_M, BN) : make_idesc_bf16(MMA_M, BN);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int item = slot; item < n_items; item += nslots) {
        const int ks = item % P.k_splits, mi = item / P.k_splits;
        const int mu = mi % m_units, split = mi / m_units;
        if (unit_skipped(mu)) continue;
        const int kb0 = P.num_k_blocks * ks / P.k_splits, kb1 = P.num_k_blocks * (ks + 1) / P.k_splits;
        const int64_t t0 = P.num_n_tiles * split / P.splits, t1 = P.num_n_tiles * (split + 1) / P.splits;
        for (int64_t t = t0; t < t1; ++t) {
          mbar_wait(smem_u32(&tail->tmem_empty[as]), aphase ^ 1);
          tc_fence_after();
#ifdef FRX_TRACE
          if (P.trace && bid == 0 && (t - t0) < 256) P.trace[(t - t0) * 4 + 0] = clock64();
#endif
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(smem_u32(&tail->full[stage]), phase);
            tc_fence_after();
            const uint64_t da = make_sw128_kmajor_desc(smem_a + stage * A_STAGE_BYTES);
            const uint64_t db = make_sw128_kmajor_desc(smem_b + stage * B_STAGE);
#pragma unroll
            for (int k = 0; k < MMAS_PER_KBLOCK; ++k) {
              // advance 16 bf16 / 8 tf32 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
              const uint32_t acc = (kb > kb0 || k > 0);
              if (PAIR) {
                if (TF32) umma_tf32_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, acc);
                else umma_bf16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, acc);
              } else {
                if (TF32) umma_tf32_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, acc);
                else umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, acc);
              }
            }
            // frees the smem slot (in both CTAs of a pair) when the MMAs retire
            if (PAIR) umma_commit_pair(smem_u32(&tail->empty[stage]), 3);
            else if (CL > 1) umma_commit_multicast(smem_u32(&tail->empty[stage]), CL_MASK);
            else umma_commit(smem_u32(&tail->empty[stage]));
            if (kb == kb1 - 1) {
              if (PAIR) umma_commit_pair(smem_u32(&tail->tmem_full[as]), 3); else umma_commit(smem_u32(&tail->tmem_full[as]));
#ifdef FRX_TRACE
              if (P.trace && bid == 0 && (t - t0) < 256) P.trace[(t - t0) * 4 + 1] = clock64();
#endif
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++as == 2) { as = 0; aphase ^= 1; }
        }
      }
    }
  } else if (warp < EPI_WARP0) {
    setmaxnreg_dec<LEAN_REGS>();                           // warps 2 (TMEM alloc / dealloc) and 3: same warp group as 0 and 1
  } else {
    // =========================== epilogue ===========================
    setmaxnreg_inc<EPI_REGS>();
    // 8 warps: lane quarter q = warp % 4 (TMEM lanes 32q..32q+31 = brand rows), column half h.
    const int q = warp & 3;
    const int h = (warp - EPI_WARP0) >> 2;
    const int ew = warp - EPI_WARP0;
    const int row_in_tile = q * 32 + lane;
    constexpr int CHUNKS = BN / HV / 32;         // chunks of 32 columns per warp per tile: 8 (4 warps) or 4 (8 warps)
    uint32_t* hist = tail->scratch[ew];
    float* stage = reinterpret_cast<float*>(tail->scratch[ew]);
    int as = 0; uint32_t aphase = 0;
#ifdef FRX_TRACE
    if (P.cta_trace && ew == 0 && lane == 0) {
      uint32_t smid; unsigned long long gt;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      P.cta_trace[bid * 4 + 0] = smid;
      P.cta_trace[bid * 4 + 1] = (long long)gt;
    }
#endif
    for (int item = slot; item < n_items; item += nslots) {
      const int ks = item % P.k_splits, mi = item / P.k_splits;
      const int mu = mi % m_units, split = mi / m_units;
      if (unit_skipped(mu)) continue;
      const int m_tile = mu * UNIT + (int)crank;              // this CTA's 128 accumulator lanes = these brand rows
      const int64_t t0 = P.num_n_tiles * split / P.splits, t1 = P.num_n_tiles * (split + 1) / P.splits;
      const int row = m_tile * BM + row_in_tile;
      const bool row_ok = row < P.nb;
      // dense output: rows of this warp inside the matrix, and whether 16-byte vector stores of row segments are aligned
      const int warp_rows = min(max(P.nb - (m_tile * BM + q * 32), 0), 32);
      const bool dense_vec_ok = P.dense != nullptr && (P.ld_dense & 3) == 0 && (P.partial_stride & 3) == 0 &&
                                (reinterpret_cast<uintptr_t>(P.dense) & 15) == 0;

      // per-row state
      float thr = row_ok ? -INFINITY : INFINITY;     // TOPK: append threshold
      int cnt = 0;                                   // TOPK: candidates buffered
      // candidate lists of this (split, m-tile, column half)
      const size_t part = ((((size_t)split * P.num_m_tiles + m_tile) * P.k_splits + ks) * HV + h) * BM;
      unsigned long long* rowbuf = nullptr;
      float ts = 0.f; int32_t ti = -1; unsigned long long ccount = 0;
      if (MODE == MODE_TOPK) rowbuf = P.part_keys + (part + row_in_tile) * P.cap;
      uint32_t hbase = 0;                            // TOPK: origin of this row's histogram bins (0 = off)
      uint32_t* hrow = nullptr;
      if (MODE == MODE_TOPK && P.row_hist != nullptr && row_ok) {
        hbase = __ldg(P.row_base + row);
        hrow = P.row_hist + (size_t)row * HIST_BINS;
      }
      if (MODE == MODE_COUNT && row_ok) { ts = P.thr_score[row]; ti = P.thr_index[row]; }

      for (int64_t t = t0; t < t1; ++t) {
        const int64_t col0 = t * BN + h * (BN / HV);
        // issued before the accumulator wait so that their latency is hidden: the labels of this warp's
        // 128 columns and the row's global threshold (best lower bound published by any CTA so far)
        int lab[CHUNKS];
        if (MODE == MODE_TOPK) {
          if (P.labels != nullptr) {
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
              const int64_t j = col0 + c * 32 + lane;
              lab[c] = j < P.n_posts ? __ldg(P.labels + j) : -1;
            }
          }
          if (row_ok) thr = fmaxf(thr, ordered_to_score(__ldcg(P.row_thr + row)));
          // every REFINE_EVERY tiles (the two column halves alternate): re-derive the threshold from what ALL CTAs
          // of this row have appended so far and publish it
          if (P.row_hist != nullptr && ((int)(t - t0) % P.refine_every) == (h ? P.refine_every / 2 : 0) && hbase != 0u) {
            const int edge = hist_edge(hrow, (uint32_t)P.k);
            if (edge > 0) {
              const uint32_t eo = hbase + ((uint32_t)edge << HIST_SHIFT);
              const float nt = ordered_to_score(eo);
              if (nt == nt && nt > thr) {
                thr = nt;
                atomicMax(P.row_thr + row, eo);
              }
            }
          }
        }
        mbar_wait(smem_u32(&tail->tmem_full[as]), aphase);
        tc_fence_after();
#ifdef FRX_TRACE
        if (P.cta_trace && ew == 0 && lane == 0 && item == slot && (t - t0) == 50) {
          unsigned long long gt;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
          P.cta_trace[bid * 4 + 2] = (long long)gt;
        }
#endif
#ifdef FRX_TRACE
        if (P.trace && bid == 0 && ew == 0 && lane == 0 && (t - t0) < 256) P.trace[(t - t0) * 4 + 2] = clock64();
#endif
#pragma unroll 1
        for (int c = 0; c < CHUNKS; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + h * (BN / HV) + c * 32), v);
          tmem_ld_wait();
          const int64_t cbase = col0 + c * 32;
          const int64_t rem = P.n_posts - cbase;
          const int nvalid = rem >= 32 ? 32 : (rem > 0 ? (int)rem : 0);
          if (nvalid == 0) continue;                 // warp-uniform

          if (MODE == MODE_TOPK) {
            // the caller also wants the score matrix (exact-AUC sweep): same accumulators, written on the way
            if (P.dense != nullptr) {
              float* wbase = P.dense + (int64_t)(m_tile * BM + q * 32) * P.ld_dense + cbase;
              if (nvalid == 32 && dense_vec_ok)
                store_dense_chunk_staged<true>(wbase, P.ld_dense, v, 1.0f, stage, lane, warp_rows);
              else if (row_ok)
                store_dense_chunk<true>(wbase + (int64_t)lane * P.ld_dense, v, nvalid, 1.0f);
            }
            if (nvalid < 32) {                       // last tile only: out-of-range columns can never qualify
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i >= nvalid) v[i] = 0x7FC00000u;   // NaN: fails every >=
            }
            if (P.labels != nullptr) {
              // S[label[j], j]: lane i holds the label of column cbase+i; the owner row is a lane of this warp
              // iff label - (m_tile*128 + q*32) is in [0, 32).  ~1 column per chunk qualifies: loop over the set
              // bits, the owner lane picks its register with a predicated select chain (no per-column branch).
              int lbl = -1;
#pragma unroll
              for (int cc = 0; cc < CHUNKS; ++cc) if (cc == c) lbl = lab[cc];
              int tgt = -1;
              if (lbl >= 0 && lbl < P.nb) tgt = lbl - (m_tile * BM + q * 32);   // out-of-range labels keep NaN
              uint32_t hit = __ballot_sync(0xffffffffu, tgt >= 0 && tgt < 32);
              while (hit) {
                const int i = __ffs(hit) - 1;
                hit &= hit - 1;
                const int owner = __shfl_sync(0xffffffffu, tgt, i);
                if (lane == owner) {
                  uint32_t val = 0;
#pragma unroll
                  for (int j = 0; j < 32; ++j) val = (j == i) ? v[j] : val;
                  P.pos_score[cbase + i] = __uint_as_float(val);
                }
              }
            }
            // per-lane qualification mask (branch-free), then ONE warp reduction tells which columns have a
            // qualifying score in any row; appends are visited in groups of 4 columns
            uint32_t hm = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) hm |= (__uint_as_float(v[i]) >= thr ? 1u : 0u) << i;
            const uint32_t colmask = __reduce_or_sync(0xffffffffu, hm);
            if (colmask) {
              const uint32_t gbase = (uint32_t)(P.index_base + cbase);
#pragma unroll
              for (int g4 = 0; g4 < 8; ++g4) {
                if (colmask & (0xFu << (4 * g4))) {  // warp-uniform
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const int i = 4 * g4 + j;
                    if ((hm >> i) & 1u) {
                      const uint32_t so = score_to_ordered(__uint_as_float(v[i]));
                      rowbuf[cnt] = ((unsigned long long)so << 32) | (unsigned long long)(0xFFFFFFFFu - (gbase + i));
                      ++cnt;
                      if (hbase != 0u) {
                        const uint32_t bin = (so - hbase) >> HIST_SHIFT;     // so >= ordered(thr) >= hbase
                        atomicAdd(hrow + (bin < (uint32_t)(HIST_BINS - 1) ? bin : (uint32_t)(HIST_BINS - 1)), 1u);
                      }
                    }
                  }
                }
              }
              // keep room for the next 32-column chunk
              uint32_t full = __ballot_sync(0xffffffffu, cnt > P.cap - 32);
              if (full) {
                __syncwarp();
                while (full) {
                  const int l = __ffs(full) - 1;
                  full &= full - 1;
                  const int n = __shfl_sync(0xffffffffu, cnt, l);
                  unsigned long long* bb = P.part_keys + (part + q * 32 + l) * P.cap;
                  float nthr;
                  const int kept = select_dispatch(P.cap, bb, n, P.k, P.keep_limit, false, hist, &nthr);
                  if (lane == l) {
                    cnt = kept;
                    thr = fmaxf(thr, nthr);
                    atomicMax(P.row_thr + row, score_to_ordered(thr));   // publish: valid for every CTA of this row
                  }
                }
                __syncwarp();
              }
            }
          } else if (MODE == MODE_DENSE) {
            float* wbase = P.dense + (int64_t)ks * P.partial_stride + (int64_t)(m_tile * BM + q * 32) * P.ld_dense + cbase;
            if (nvalid == 32 && dense_vec_ok)
              store_dense_chunk_staged<false>(wbase, P.ld_dense, v, P.dense_scale, stage, lane, warp_rows);
            else if (row_ok)
              store_dense_chunk<false>(wbase + (int64_t)lane * P.ld_dense, v, nvalid, P.dense_scale);
          } else {   // MODE_COUNT
            if (ti >= 0) {
              const int64_t gbase = P.index_base + cbase;
              int local = 0;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float s = __uint_as_float(v[i]);
                const bool before = (s > ts) || (s == ts && (gbase + i) < (int64_t)ti);
                local += (i < nvalid && before) ? 1 : 0;
              }
              ccount += (unsigned long long)local;
            }
          }
        }
        // release this accumulator stage to the MMA warp
        tc_fence_before();
        __syncwarp();
#ifdef FRX_TRACE
        if (P.trace && bid == 0 && ew == 0 && lane == 0 && (t - t0) < 256) P.trace[(t - t0) * 4 + 3] = clock64();
#endif
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&tail->tmem_empty[as]), 0));   // the leader's barrier
          else mbar_arrive(smem_u32(&tail->tmem_empty[as]));
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }

      // ---- end of item ----
      if (MODE == MODE_TOPK) {
        uint32_t over = __ballot_sync(0xffffffffu, cnt > P.k);
        __syncwarp();
        while (over) {
          const int l = __ffs(over) - 1;
          over &= over - 1;
          const int n = __shfl_sync(0xffffffffu, cnt, l);
          unsigned long long* b = P.part_keys + (part + q * 32 + l) * P.cap;
          float nthr;
          const int kept = select_dispatch(P.cap, b, n, P.k, P.k, true, hist, &nthr);
          if (lane == l) { cnt = kept; atomicMax(P.row_thr + row, score_to_ordered(fmaxf(thr, nthr))); }
        }
        P.part_cnt[part + row_in_tile] = cnt;
      } else if (MODE == MODE_COUNT) {
        if (row_ok && ti >= 0 && ccount) atomicAdd(P.count_out + row, ccount);
      }
    }
  }

  // =========================== teardown ===========================
#ifdef FRX_TRACE
  if (P.cta_trace && tid == EPI_WARP0 * 32) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    P.cta_trace[bid * 4 + 3] = (long long)gt;
  }
#endif
  tc_fence_before();
  if (CLUSTERED) cluster_sync_all(); else __syncthreads();  // no CTA of a cluster exits while another may still signal it
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int MODE, bool TF32>
__global__ void __maxnreg__(KERNEL_REGS)
score_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ ScoreParams P) {
  score_body<MODE, TF32, false, 1, NUM_EPI_WARPS>(tmap_a, tmap_b, P);
}

// The wide CTA (8 epilogue warps, 384 threads): fused top-k at large k.
template <int MODE, bool TF32>
__global__ void __maxnreg__(WIDE_REGS)
score_kernel_wide(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ ScoreParams P) {
  score_body<MODE, TF32, false, 1, WIDE_EPI_WARPS>(tmap_a, tmap_b, P);
}

// The same kernel in clusters of CL CTAs that share every post tile through TMA multicast (cluster size given at launch).
template <int MODE, bool TF32, int CL>
__global__ void __maxnreg__(KERNEL_REGS)
score_kernel_mc(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ ScoreParams P) {
  score_body<MODE, TF32, false, CL, NUM_EPI_WARPS>(tmap_a, tmap_b, P);
}

// The same kernel on CTA pairs: clusters of two CTAs (the two SMs of a TPC), tcgen05 cta_group::2.
template <int MODE, bool TF32>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(KERNEL_REGS)
score_kernel_pair(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ ScoreParams P) {
  score_body<MODE, TF32, true, 1, NUM_EPI_WARPS>(tmap_a, tmap_b, P);
}

// ---------------------------------------------------------------------------------------------
// Block-wide bitonic sort (descending) of n_pow2 u64 keys in shared memory.
// ---------------------------------------------------------------------------------------------
__device__ void block_bitonic_desc(unsigned long long* keys, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));   // stride is a power of two
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        unsigned long long a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// One block per brand: gather the per-split candidate lists of this row, select + sort, emit the top-k.
// Candidates are staged in shared memory when they fit (`smem_keys` entries incl. room for the k survivors);
// otherwise (large k x many lists) the radix-select passes stream them from global memory through a
// flattened index (binary search over the list offsets), 8 independent loads in flight per thread.
__global__ void __launch_bounds__(256) merge_partials_kernel(const unsigned long long* __restrict__ part_keys,
                                                             const int* __restrict__ part_cnt,
                                                             const uint32_t* __restrict__ row_thr, int num_m_tiles,
                                                             int splits, int halves, int cap, int k, int smem_keys,
                                                             float* __restrict__ out_s, int32_t* __restrict__ out_i) {
  extern __shared__ unsigned long long skeys[];
  __shared__ int offs[2049];
  __shared__ int kept;
  __shared__ uint32_t hist[256];
  __shared__ int sel[3];
  const int b = blockIdx.x, m_tile = b / BM, r = b % BM;
  const int lists = splits * halves;            // (split, column range) -> list (split*num_m_tiles + m_tile)*halves + range
  auto list_ptr = [&](int s) {
    return part_keys + ((((size_t)(s / halves) * num_m_tiles + m_tile) * halves + (s % halves)) * BM + r) * cap;
  };
  // list sizes: loaded in parallel (one L2 round trip), then a serial prefix over shared memory
  for (int s = threadIdx.x; s < lists; s += blockDim.x)
    offs[s + 1] = part_cnt[(((size_t)(s / halves) * num_m_tiles + m_tile) * halves + (s % halves)) * BM + r];
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    offs[0] = 0;
    for (int s = 0; s < lists; ++s) { acc += offs[s + 1]; offs[s + 1] = acc; }
    kept = 0;
  }
  __syncthreads();
  int total = offs[lists];
  bool staged = total + (total > k ? k : 0) <= smem_keys;
  // flattened element g -> key (global path)
  auto fetch = [&](int g) -> unsigned long long {
    int lo = 0, hi = lists;                      // largest s with offs[s] <= g
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offs[mid] <= g) lo = mid; else hi = mid; }
    return __ldcg(list_ptr(lo) + (g - offs[lo]));
  };
  if (staged) {
    // flattened gather: every candidate of every list is one independent load (a per-list loop would serialise
    // dozens of L2 round trips over lists that hold a handful of entries each)
    for (int g = threadIdx.x; g < total; g += blockDim.x) skeys[g] = fetch(g);
    __syncthreads();
  } else {
    // Too many candidates for shared memory: most of them were appended under an early, loose threshold.  Keep only
    // those that reach the row's FINAL published threshold (a lower bound of the k-th best, so no top-k entry is
    // lost); when the survivors fit, carry on in shared memory, else stream everything from global memory.
    const unsigned long long min_key = (unsigned long long)row_thr[b] << 32;
    const int room = smem_keys - k;
    for (int g0 = threadIdx.x; g0 < total; g0 += 8 * 256) {
      unsigned long long kk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int g = g0 + j * 256; kk[j] = g < total ? fetch(g) : 0ull; }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (g0 + j * 256 < total && kk[j] >= min_key) {
          const int slot = atomicAdd(&kept, 1);
          if (slot < room) skeys[slot] = kk[j];
        }
      }
    }
    __syncthreads();
    const int nf = kept;
    __syncthreads();
    if (threadIdx.x == 0) kept = 0;
    __syncthreads();
    if (nf <= room) { total = nf; staged = true; }
  }
  if (total > k) {
    // block-wide MSB-first radix select of the k-th best key (exact; keys are unique)
    unsigned long long prefix = 0;
    int need = k, shift = 56;
    for (int pass = 0; pass < 8; ++pass, shift -= 8) {
      if (threadIdx.x < 256) hist[threadIdx.x] = 0;
      __syncthreads();
      if (staged) {
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
          const unsigned long long key = skeys[i];
          if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
        }
      } else {
        for (int g0 = threadIdx.x; g0 < total; g0 += 8 * 256) {
          unsigned long long kk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { const int g = g0 + j * 256; kk[j] = g < total ? fetch(g) : 0ull; }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (g0 + j * 256 < total && (pass == 0 || (kk[j] >> (shift + 8)) == prefix))
              atomicAdd(&hist[(uint32_t)(kk[j] >> shift) & 255u], 1u);
        }
      }
      __syncthreads();
      if (threadIdx.x < 32) {
        int digit, above, bucket;
        select_digit(hist, threadIdx.x, need, digit, above, bucket);
        if (threadIdx.x == 0) { sel[0] = digit; sel[1] = above; sel[2] = bucket; }
      }
      __syncthreads();
      prefix = (prefix << 8) | (unsigned long long)sel[0];
      need -= sel[1];
      const int bucket = sel[2];
      __syncthreads();
      if (bucket == 1 || pass == 7) break;        // a singleton bucket pins the k-th key
    }
    const unsigned long long thr_key = prefix << shift;
    // the k survivors: staged -> tail region of skeys then moved to the front; global -> straight into skeys
    unsigned long long* dst = staged ? skeys + total : skeys;
    if (staged) {
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const unsigned long long key = skeys[i];
        if (key >= thr_key) dst[atomicAdd(&kept, 1)] = key;
      }
    } else {
      for (int g0 = threadIdx.x; g0 < total; g0 += 8 * 256) {
        unsigned long long kk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int g = g0 + j * 256; kk[j] = g < total ? fetch(g) : 0ull; }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (g0 + j * 256 < total && kk[j] >= thr_key) dst[atomicAdd(&kept, 1)] = kk[j];
      }
    }
    __syncthreads();
    const int nk = kept;
    if (staged) {
      for (int i = threadIdx.x; i < nk; i += blockDim.x) { const unsigned long long key = dst[i]; skeys[i] = key; }
      __syncthreads();                            // dst starts at total >= nk: reads and writes never alias
    }
    total = nk;
  } else if (!staged) {
    for (int g = threadIdx.x; g < total; g += blockDim.x) skeys[g] = fetch(g);
    __syncthreads();
  }
  const int np2 = next_pow2(total > 1 ? total : 2);
  for (int i = total + threadIdx.x; i < np2; i += blockDim.x) skeys[i] = 0ull;
  block_bitonic_desc(skeys, np2);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    if (i < total) {
      out_s[(size_t)b * k + i] = key_score(skeys[i]);
      out_i[(size_t)b * k + i] = (int32_t)key_index(skeys[i]);
    } else {
      out_s[(size_t)b * k + i] = -INFINITY;
      out_i[(size_t)b * k + i] = -1;
    }
  }
}

// Sample pass, dense flavour: one block per brand row finds the k-th largest of the row's n sample scores
// (MSB-first 8-bit radix select on order-preserving u32 keys held in shared memory) and writes it to row_thr.
__global__ void __launch_bounds__(256) row_kth_kernel(const float* __restrict__ dense, int64_t ld, int n, int k,
                                                      uint32_t* __restrict__ row_thr) {
  extern __shared__ uint32_t sk32[];
  __shared__ uint32_t hist[256];
  __shared__ int sel[3];
  const int b = blockIdx.x;
  const float* row = dense + (int64_t)b * ld;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float s = row[i];
    sk32[i] = (s == s) ? score_to_ordered(s) : 0u;      // NaN ranks below everything
  }
  uint32_t prefix = 0;
  int need = k < n ? k : n, shift = 24;
  for (int pass = 0; pass < 4; ++pass, shift -= 8) {
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    const int n_round = (n + 255) & ~255;                 // whole warps stay converged for match_any
    for (int i = threadIdx.x; i < n_round; i += blockDim.x) {
      const uint32_t key = i < n ? sk32[i] : 0u;
      const bool on = i < n && (pass == 0 || (key >> (shift + 8)) == prefix);
      // scores cluster in a few digits (same sign/exponent): one shared-memory atomic per distinct digit per warp
      const uint32_t digit = on ? ((key >> shift) & 255u) : 0xFFFFFFFFu;
      if (pass == 0) {
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        if (on && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
      } else if (on) {
        atomicAdd(&hist[digit], 1u);              // later digits are mantissa bits: spread out, few lanes active
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      int digit, above, bucket;
      select_digit(hist, threadIdx.x, need, digit, above, bucket);
      if (threadIdx.x == 0) { sel[0] = digit; sel[1] = above; }
    }
    __syncthreads();
    prefix = (prefix << 8) | (uint32_t)sel[0];
    need -= sel[1];
    __syncthreads();
  }
  if (threadIdx.x == 0) row_thr[b] = prefix;            // exact k-th largest score of the sample
}

// row_thr[b] = ordered(k-th best score of the sample pass) -- a valid lower bound of the row's global k-th best
// because the sample is a subset of the posts.  Rows whose sample list is short keep 0 (= no threshold).
__global__ void seed_threshold_kernel(const float* __restrict__ topk_scores, int nb, int k, uint32_t* __restrict__ row_thr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const float s = topk_scores[(size_t)b * k + (k - 1)];
  row_thr[b] = (s == s && s > -INFINITY) ? score_to_ordered(s) : 0u;
}

// Merge g lists [g, nb, k_in] of (score, index) into [nb, k_out] (multi-GPU exchange step).  Every input list is
// SORTED under the total order with its padding (index < 0) at the tail -- it is a top-k list of frx_score_topk -- so the
// global rank of an entry is its position in its own list plus, for every other list, the number of entries that precede
// it there (keys are unique: distinct posts), found by binary search.  No sorting network, no block barrier per stage:
// one independent chain of g - 1 searches per entry (8 lists of 1000: 7.7 ms -> see profiles for the bitonic version).
__global__ void __launch_bounds__(256) merge_lists_kernel(const float* __restrict__ in_s, const int32_t* __restrict__ in_i,
                                                          int g, int nb, int k_in, int64_t shard_stride,
                                                          float* __restrict__ out_s, int32_t* __restrict__ out_i, int k_out) {
  extern __shared__ unsigned long long skeys[];           // [g * k_in] keys, then [k_out] merged keys
  unsigned long long* merged = skeys + (size_t)g * k_in;
  const int b = blockIdx.x;
  const int total = g * k_in;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int gi = i / k_in, ki = i % k_in;
    const size_t src = (size_t)gi * shard_stride + (size_t)b * k_in + ki;
    const int32_t idx = in_i[src];
    skeys[i] = idx >= 0 ? make_key(in_s[src], (uint32_t)idx) : 0ull;   // 0 = padding: below every real key
  }
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) merged[i] = 0ull;
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const unsigned long long key = skeys[i];
    if (key == 0ull) continue;
    const int gi = i / k_in;
    int rank = i - gi * k_in;
    for (int o = 0; o < g && rank < k_out; ++o) {
      if (o == gi) continue;
      const unsigned long long* lst = skeys + (size_t)o * k_in;
      int lo = 0, hi = k_in;                                // number of entries of list o that precede `key`
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (lst[mid] > key) lo = mid + 1; else hi = mid; }
      rank += lo;
    }
    if (rank < k_out) merged[rank] = key;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
    const unsigned long long key = merged[i];
    if (key != 0ull) {
      out_s[(size_t)b * k_out + i] = key_score(key);
      out_i[(size_t)b * k_out + i] = (int32_t)key_index(key);
    } else {
      out_s[(size_t)b * k_out + i] = -INFINITY;
      out_i[(size_t)b * k_out + i] = -1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// [rows, d] bf16 row-major with leading dimension ld -> 2-D tensor map, box = 64 (K) x box_rows,
// SWIZZLE_128B, out-of-bounds elements (row tails, K tail) read as zero.
static int make_operand_map(CUtensorMap* map, const void* ptr, int64_t rows, int d, int64_t ld, int box_rows, bool tf32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return FRX_E_DEVICE; }
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  const int esize = tf32 ? 4 : 2;
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)(BK_BYTES / esize), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return FRX_E_CUDA; }
  return FRX_OK;
}

// Two variants of the same kernel: one CTA per SM (default), or CTA pairs (cta_group::2, one 256 x 256 MMA tile per
// pair).  Measured on B200 under its 1 kW cap: the pair variant needs ~6 % fewer cycles per tile at
// D = 3072 (a third less L2 -> shared-memory operand traffic, deeper ring) but the cross-SM operand reads cost power, the
// SM clock settles ~7 % lower and wall-clock time is equal or slightly worse -- so it is opt-in: FRX_PAIR=1 or
// frx_set_cta_pairs(1).
static int g_pair_mode = -1;                      // -1: take the FRX_PAIR environment variable (default 0)
static bool use_pair() {
  static const int env = getenv("FRX_PAIR") ? atoi(getenv("FRX_PAIR")) : 0;
  const int v = g_pair_mode >= 0 ? g_pair_mode : env;
  return v != 0 && num_sms() % 2 == 0;
}

// Third variant: clusters of CL CTAs sharing every post tile through TMA multicast (score_kernel_mc).  FRX_CLUSTER = 2 / 4 / 8
// or frx_set_cluster(); bf16 operands only; 0 / 1 = off.
static int g_cluster_mode = -1;
static int use_cluster() {
  static const int env = getenv("FRX_CLUSTER") ? atoi(getenv("FRX_CLUSTER")) : 0;
  const int v = g_cluster_mode >= 0 ? g_cluster_mode : env;
  return (v == 2 || v == 4 || v == 8) ? v : 1;
}

struct Plan {
  int cluster;        // > 1: multicast clusters of this many CTAs (one scheduling unit = cluster x 128 brand rows)
  bool wide;          // the 8-epilogue-warp CTA (fused top-k with k > WIDE_K on the default variant)
  int halves;         // candidate lists per (item, brand row): 2 on the wide CTA, else 1
  bool pair;          // one scheduling unit = a CTA pair working on 256 brand rows
  int num_m_tiles;    // 128-row m-tiles (padded to an even count for pairs: candidate lists are addressed by m-tile)
  int m_units, slots; // schedulable m-units (m-tiles or pairs of them) and concurrently resident units
  int splits, cap, keep_limit, grid;
  int64_t num_n_tiles;
  size_t keys_bytes, cnt_bytes, thr_bytes;
};

static int max_active_clusters(int cl);

static Plan make_plan(int nb, int64_t n_posts, int k, int mode, bool tf32 = false) {
  Plan p{};
  p.cluster = tf32 ? 1 : use_cluster();
  p.pair = p.cluster == 1 && use_pair();
  const int real_m_tiles = (nb + BM - 1) / BM;
  const int unit = p.pair ? 2 : p.cluster;
  p.num_m_tiles = (real_m_tiles + unit - 1) / unit * unit;
  p.m_units = p.num_m_tiles / unit;
  p.slots = p.pair ? num_sms() / 2 : (p.cluster > 1 ? max_active_clusters(p.cluster) : num_sms());
  if (p.slots < 1) { p.cluster = 1; p.pair = false; p.num_m_tiles = real_m_tiles; p.m_units = real_m_tiles; p.slots = num_sms(); }
  const int sms = p.slots;
  p.num_n_tiles = (n_posts + BN - 1) / BN;
  int64_t smax = p.num_n_tiles;
  if (smax > 1024) smax = 1024;
  if (smax < 1) smax = 1;
  // pick the split count that fills whole waves of `sms` CTAs; prefer fewer, longer items
  int best = 1; double best_eff = -1.0;
  for (int s = 1; s <= (int)smax; ++s) {
    const long items = (long)p.m_units * s;
    const long waves = (items + sms - 1) / sms;
    double eff = (double)items / (double)(waves * sms);
    if (waves > 8 && s > 1) break;               // long enough; more splits only add merge work
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  p.splits = best;
  const long units = (long)p.m_units * p.splits;
  p.grid = (int)(units < sms ? units : sms) * (p.pair ? 2 : p.cluster);
  const long items = (long)p.num_m_tiles * p.splits;             // candidate lists exist per (split, m-tile)
  int cap = 1024;
  while (cap < 4 * k) cap <<= 1;
  p.cap = cap;
  p.keep_limit = k + k / 4;
  p.wide = mode == MODE_TOPK && !tf32 && p.cluster == 1 && !p.pair && k > WIDE_K;
  p.halves = p.wide ? 2 : 1;
  p.keys_bytes = (size_t)items * p.halves * BM * cap * sizeof(unsigned long long);
  p.cnt_bytes = (((size_t)items * p.halves * BM * sizeof(int)) + 255) & ~(size_t)255;
  p.thr_bytes = (((size_t)nb * sizeof(uint32_t)) + 255) & ~(size_t)255;
  return p;
}

// ---- measurement hook: CUDA events around each score_kernel launch --------------------------
#ifdef FRX_TRACE
static long long* g_trace = nullptr;     // debug builds only (tools/gpu_trace_pipeline.py)
static long long* g_cta_trace = nullptr; // debug builds only (tools/gpu_trace_ctas.py)
#endif

struct Probe {
  bool on = false;
  int n = 0;
  cudaEvent_t beg[4096], end[4096];
  bool made[4096] = {};
};
static Probe g_probe;

// Workspace layout of frx_score_topk: [part_cnt | row_thr | part_keys], each region sized for the larger of
// the main pass and the sample pass (which reuses them).
struct TopkLayout {
  Plan main, sample;
  bool has_sample, sample_dense;
  int64_t n_s, stride;
  size_t cnt_bytes, thr_bytes, keys_bytes, dense_bytes, hist_bytes;
  // [part_cnt | row_thr | row_base | row_hist | part_keys | sample dense tile]
  size_t total() const { return cnt_bytes + 2 * thr_bytes + hist_bytes + keys_bytes + dense_bytes + 256; }
};

static TopkLayout make_topk_layout(int nb, int64_t n_posts, int k, bool tf32 = false) {
  TopkLayout L{};
  L.main = make_plan(nb, n_posts, k, MODE_TOPK, tf32);
  L.cnt_bytes = L.main.cnt_bytes;
  L.keys_bytes = L.main.keys_bytes;
  L.thr_bytes = L.main.thr_bytes;
  L.hist_bytes = (size_t)nb * HIST_BINS * sizeof(uint32_t);
  L.has_sample = n_posts >= kSampleMinPosts;
  if (L.has_sample) {
    // The sample only has to put the seeded threshold within the histogram's range (two octaves) of the final
    // k-th best score; the refinement does the rest.  16 k posts ... 4096 at least, 32768 at most, <= 1/16 of the posts.
    int64_t n_s = 16 * (int64_t)k;
    n_s = n_s < 4096 ? 4096 : (n_s > 32768 ? 32768 : n_s);
    if (n_s > n_posts / 16) n_s = n_posts / 16;
    if (n_s < 4 * (int64_t)k) n_s = 4 * (int64_t)k;
    n_s = (n_s + BN - 1) / BN * BN;
    L.n_s = n_s;
    L.stride = n_posts / n_s;
    // small enough: dense sample tile + per-row k-th select (cheapest); else the fused top-k kernel on the sample
    L.sample_dense = (size_t)nb * (size_t)n_s * sizeof(float) <= ((size_t)1 << 30) && n_s <= 32768;
    if (L.sample_dense) {
      L.sample = make_plan(nb, n_s, 1, MODE_DENSE, tf32);
      L.dense_bytes = (((size_t)nb * (size_t)n_s * sizeof(float)) + 255) & ~(size_t)255;
    } else {
      L.sample = make_plan(nb, n_s, k, MODE_TOPK, tf32);
      if (L.sample.cnt_bytes > L.cnt_bytes) L.cnt_bytes = L.sample.cnt_bytes;
      if (L.sample.keys_bytes > L.keys_bytes) L.keys_bytes = L.sample.keys_bytes;
    }
  }
  return L;
}

template <int MODE, int CL>
static int launch_mc(int grid, const CUtensorMap& ma, const CUtensorMap& mb, const ScoreParams& P, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    FRX_CUDA(cudaFuncSetAttribute(score_kernel_mc<MODE, false, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  FRX_CUDA(cudaLaunchKernelEx(&cfg, score_kernel_mc<MODE, false, CL>, ma, mb, P));
  return FRX_OK;
}

// resident clusters of `cl` multicast CTAs on this device (all modes have the same footprint)
static int max_active_clusters(int cl) {
  static int cached[9] = {0};
  if (cl < 2 || cl > 8) return 0;
  if (cached[cl]) return cached[cl] > 0 ? cached[cl] : 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(num_sms() / cl * cl));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = (unsigned)cl; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaErrorInvalidValue;
  if (cl == 2) { cudaFuncSetAttribute(score_kernel_mc<MODE_TOPK, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
                 e = cudaOccupancyMaxActiveClusters(&n, score_kernel_mc<MODE_TOPK, false, 2>, &cfg); }
  if (cl == 4) { cudaFuncSetAttribute(score_kernel_mc<MODE_TOPK, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
                 e = cudaOccupancyMaxActiveClusters(&n, score_kernel_mc<MODE_TOPK, false, 4>, &cfg); }
  if (cl == 8) { cudaFuncSetAttribute(score_kernel_mc<MODE_TOPK, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
                 e = cudaOccupancyMaxActiveClusters(&n, score_kernel_mc<MODE_TOPK, false, 8>, &cfg); }
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  cached[cl] = n > 0 ? n : -1;
  return n;
}

template <int MODE, bool TF32>
static int launch_score(const void* a, int64_t ld_a, const void* b, int64_t ld_b, int nb, int64_t n_posts, int d,
                        const Plan& plan, ScoreParams& P, cudaStream_t st, bool allow_probe = true) {
  CUtensorMap ma, mb;
  int rc = make_operand_map(&ma, a, nb, d, ld_a, BM, TF32);
  if (rc) return rc;
  // a pair's CTAs load half a B tile each; a multicast cluster's CTAs 1 / CL of it
  rc = make_operand_map(&mb, b, n_posts, d, ld_b, plan.pair ? BN / 2 : BN / plan.cluster, TF32);
  if (rc) return rc;
  P.nb = nb;
  P.n_posts = n_posts;
  constexpr int BK = TF32 ? 32 : 64;
  P.num_k_blocks = (d + BK - 1) / BK;
  P.num_m_tiles = plan.num_m_tiles;
  P.num_n_tiles = plan.num_n_tiles;
  P.splits = plan.splits;
  if (P.k_splits < 1) P.k_splits = 1;
#ifdef FRX_TRACE
  P.trace = (MODE == MODE_TOPK && allow_probe) ? g_trace : nullptr;
  P.cta_trace = (MODE == MODE_TOPK && allow_probe) ? g_cta_trace : nullptr;
#endif
  if (plan.cluster > 1 && !TF32) { /* attribute set in launch_mc */ }
  else if (plan.pair)
    FRX_CUDA(cudaFuncSetAttribute(score_kernel_pair<MODE, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  else
    FRX_CUDA(cudaFuncSetAttribute(score_kernel<MODE, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  const bool probe = allow_probe && MODE == MODE_TOPK && g_probe.on && g_probe.n < 4096;   // main fused launches only
  const int slot = g_probe.n;
  if (probe) {
    if (!g_probe.made[slot]) {
      FRX_CUDA(cudaEventCreate(&g_probe.beg[slot]));
      FRX_CUDA(cudaEventCreate(&g_probe.end[slot]));
      g_probe.made[slot] = true;
    }
    FRX_CUDA(cudaEventRecord(g_probe.beg[slot], st));
  }
  const long total_units = (long)plan.m_units * plan.splits * P.k_splits;
  const int grid = (int)(total_units < plan.slots ? total_units : plan.slots) * (plan.pair ? 2 : plan.cluster);
  if (plan.cluster > 1 && !TF32) {
    rc = plan.cluster == 2 ? launch_mc<MODE, 2>(grid, ma, mb, P, st)
       : plan.cluster == 4 ? launch_mc<MODE, 4>(grid, ma, mb, P, st) : launch_mc<MODE, 8>(grid, ma, mb, P, st);
    if (rc) return rc;
  } else if (plan.pair) score_kernel_pair<MODE, TF32><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, P);   // clusters of 2 CTAs
  else if (plan.wide) {
    if constexpr (MODE == MODE_TOPK && !TF32) {
      static bool attr = false;
      if (!attr) {
        FRX_CUDA(cudaFuncSetAttribute(score_kernel_wide<MODE, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr = true;
      }
      score_kernel_wide<MODE, TF32><<<grid, WIDE_THREADS, SMEM_BYTES, st>>>(ma, mb, P);
    } else {
      set_error("internal: the wide CTA exists for the bf16 fused top-k only");
      return FRX_E_ARG;
    }
  } else score_kernel<MODE, TF32><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, P);
  FRX_LAUNCH_CHECK();
  if (probe) {
    FRX_CUDA(cudaEventRecord(g_probe.end[slot], st));
    g_probe.n = slot + 1;
  }
  return FRX_OK;
}

static int check_operands(const char* fn, const void* a, int64_t ld_a, const void* b, int64_t ld_b, int nb,
                          int64_t n_posts, int d, int64_t index_base, bool tf32) {
  const int lda_mult = tf32 ? 4 : 8;      // row pitch must be a multiple of 16 bytes for TMA
  FRX_CHECK_ARG(a && b, "%s: NULL operand", fn);
  FRX_CHECK_ARG(nb > 0 && n_posts > 0 && d > 0, "%s: empty problem nb=%d n_posts=%lld d=%d", fn, nb, (long long)n_posts, d);
  FRX_CHECK_ARG(ld_a >= d && ld_b >= d && ld_a % lda_mult == 0 && ld_b % lda_mult == 0,
                "%s: leading dimensions must be >= d and multiples of %d", fn, lda_mult);
  FRX_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0, "%s: operands must be 16-byte aligned", fn);
  FRX_CHECK_ARG(index_base >= 0 && index_base + n_posts <= 2147483647LL, "%s: index_base + n_posts must fit int32", fn);
  int dev = 0;
  FRX_CUDA(cudaGetDevice(&dev));
  return frx_device_check(dev);
}

// out[i] = scale * sum_ks partial[ks][i], summed in k-split order (deterministic)
__global__ void reduce_ksplit_kernel(const float* __restrict__ partial, int64_t stride, int k_splits, int64_t rows,
                                     int64_t cols, int64_t ld_partial, float* __restrict__ out, int64_t ld_out, float scale) {
  const int64_t n = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    float acc = 0.f;
    for (int ks = 0; ks < k_splits; ++ks) acc += partial[(int64_t)ks * stride + r * ld_partial + c];
    out[r * ld_out + c] = acc * scale;
  }
}

template <bool TF32>
static int score_dense_impl(const void* a, int64_t ld_a, const void* b, int64_t ld_b, int nb, int64_t n_posts, int d,
                            float* dense_out, int64_t ld_dense, void* stream, float scale = 1.0f,
                            void* ksplit_ws = nullptr, size_t ksplit_ws_bytes = 0) {
  int rc = check_operands("frx_score_dense", a, ld_a, b, ld_b, nb, n_posts, d, 0, TF32);
  if (rc) return rc;
  FRX_CHECK_ARG(dense_out && ld_dense >= n_posts, "frx_score_dense: bad output");
  Plan plan = make_plan(nb, n_posts, 1, MODE_DENSE, TF32);
  ScoreParams P{};
  P.ld_dense = ld_dense;
  // Small outputs (the B x B loss tile) cover only a few tiles: split K over the idle SMs and reduce afterwards.
  const int sms = num_sms();
  const long tile_units = (long)plan.m_units * plan.num_n_tiles * (plan.pair ? 2 : 1);   // SMs busy without a K split
  const int nkb = (d + (TF32 ? 32 : 64) - 1) / (TF32 ? 32 : 64);
  int k_splits = 1;
  if (ksplit_ws != nullptr && tile_units * 2 <= sms) {
    k_splits = (int)(sms / tile_units);
    if (k_splits > nkb / 4) k_splits = nkb / 4;                 // at least 4 k-blocks per item
    const size_t per_split = (size_t)nb * (size_t)n_posts * sizeof(float);
    if (per_split > 0 && (size_t)k_splits > ksplit_ws_bytes / per_split) k_splits = (int)(ksplit_ws_bytes / per_split);
    if (k_splits < 1) k_splits = 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (k_splits == 1) {
    P.dense = dense_out;
    P.dense_scale = scale;
    return launch_score<MODE_DENSE, TF32>(a, ld_a, b, ld_b, nb, n_posts, d, plan, P, st);
  }
  plan.splits = (int)plan.num_n_tiles;                          // one n-tile per item, K split k_splits ways
  P.dense = reinterpret_cast<float*>(ksplit_ws);
  P.ld_dense = n_posts;
  P.partial_stride = (int64_t)nb * n_posts;
  P.k_splits = k_splits;
  P.dense_scale = 1.0f;
  rc = launch_score<MODE_DENSE, TF32>(a, ld_a, b, ld_b, nb, n_posts, d, plan, P, st);
  if (rc) return rc;
  const int64_t n = (int64_t)nb * n_posts;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  reduce_ksplit_kernel<<<(int)blocks, 256, 0, st>>>(P.dense, P.partial_stride, k_splits, nb, n_posts, n_posts, dense_out,
                                                    ld_dense, scale);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

template <bool TF32>
static int score_count_impl(const void* a, int64_t ld_a, const void* b, int64_t ld_b, int nb, int64_t n_posts, int d,
                            int64_t index_base, const float* thr_score, const int32_t* thr_index,
                            unsigned long long* count_out, void* stream) {
  int rc = check_operands("frx_score_count", a, ld_a, b, ld_b, nb, n_posts, d, index_base, TF32);
  if (rc) return rc;
  FRX_CHECK_ARG(thr_score && thr_index && count_out, "frx_score_count: NULL pointer");
  Plan plan = make_plan(nb, n_posts, 1, MODE_COUNT, TF32);
  ScoreParams P{};
  P.index_base = index_base;
  P.thr_score = thr_score;
  P.thr_index = thr_index;
  P.count_out = count_out;
  return launch_score<MODE_COUNT, TF32>(a, ld_a, b, ld_b, nb, n_posts, d, plan, P, (cudaStream_t)stream);
}

template <bool TF32>
static int score_topk_impl(const void* a, int64_t ld_a, const void* b, int64_t ld_b, int nb, int64_t n_posts, int d, int k,
                           const int32_t* labels, int64_t index_base, float* topk_scores, int32_t* topk_index,
                           float* pos_score, float* dense_out, int64_t ld_dense, void* workspace, size_t workspace_bytes,
                           void* stream) {
  int rc = check_operands("frx_score_topk", a, ld_a, b, ld_b, nb, n_posts, d, index_base, TF32);
  if (rc) return rc;
  FRX_CHECK_ARG(k >= 1 && k <= 1024, "frx_score_topk: k=%d outside 1..1024", k);
  FRX_CHECK_ARG(topk_scores && topk_index, "frx_score_topk: NULL output");
  FRX_CHECK_ARG((labels == nullptr) == (pos_score == nullptr), "frx_score_topk: labels and pos_score go together");
  FRX_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "frx_score_topk: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const TopkLayout L = make_topk_layout(nb, n_posts, k, TF32);
  const Plan& plan = L.main;
  const size_t need = L.total();
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("frx_score_topk: workspace %zu bytes, need %zu", workspace_bytes, need);
    return FRX_E_WORKSPACE;
  }
  ScoreParams P{};
  P.index_base = index_base;
  P.k = k;
  P.cap = plan.cap;
  P.keep_limit = plan.keep_limit;
  P.part_cnt = reinterpret_cast<int*>(workspace);
  P.row_thr = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + L.cnt_bytes);
  uint8_t* const ws8 = reinterpret_cast<uint8_t*>(workspace);
  uint32_t* const row_base = reinterpret_cast<uint32_t*>(ws8 + L.cnt_bytes + L.thr_bytes);
  uint32_t* const row_hist = reinterpret_cast<uint32_t*>(ws8 + L.cnt_bytes + 2 * L.thr_bytes);
  P.part_keys = reinterpret_cast<unsigned long long*>(ws8 + L.cnt_bytes + 2 * L.thr_bytes + L.hist_bytes);
  FRX_CUDA(cudaMemsetAsync(P.row_thr, 0, L.thr_bytes, st));        // ordered 0 = below every score
  P.labels = labels;
  P.pos_score = pos_score;
  if (pos_score) FRX_CUDA(cudaMemsetAsync(pos_score, 0xFF, (size_t)n_posts * sizeof(float), st));   // NaN
  auto run_merge = [&](const Plan& pl) -> int {
    // shared memory: every candidate + the k survivors when that fits MAX_MERGE_KEYS, else just the sort buffer
    // (>= 2k keys) and the kernel streams the candidates from global memory
    const size_t total_max = (size_t)pl.splits * pl.halves * (size_t)k;
    size_t np2 = 2;
    while (np2 < (size_t)2 * k) np2 <<= 1;
    size_t mkeys = total_max + k;
    if (mkeys > (size_t)MAX_MERGE_KEYS + 1024) mkeys = (size_t)8 * k + 1024;   // room for the threshold-filtered candidates
    if (mkeys > (size_t)MAX_MERGE_KEYS + 1024) mkeys = (size_t)MAX_MERGE_KEYS + 1024;
    if (mkeys < np2) mkeys = np2;
    const size_t msmem = mkeys * sizeof(unsigned long long);
    if (msmem > 32 * 1024)   // static smem (~10 KB) counts against the 48 KB default too
      FRX_CUDA(cudaFuncSetAttribute(merge_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    merge_partials_kernel<<<nb, 256, msmem, st>>>(P.part_keys, P.part_cnt, P.row_thr, pl.num_m_tiles, pl.splits, pl.halves, pl.cap, k,
                                                  (int)mkeys, topk_scores, topk_index);
    FRX_LAUNCH_CHECK();
    return FRX_OK;
  };
  // ---- sample pass: seed the per-row thresholds ------------------------------------------------
  // The fused kernel is first run on a strided sample of the posts (a TMA view with a larger row pitch:
  // no data is moved).  The k-th best score of the sample is a valid lower bound of the global k-th
  // best, so the main pass starts with a pass rate of ~k/n_sample instead of warming every candidate
  // list up from -inf; candidates appended per row drop by an order of magnitude.
  if (L.has_sample && L.sample_dense) {
    float* sdense = reinterpret_cast<float*>(ws8 + L.cnt_bytes + 2 * L.thr_bytes + L.hist_bytes + L.keys_bytes);
    ScoreParams S{};
    S.dense = sdense;
    S.ld_dense = L.n_s;
    S.dense_scale = 1.0f;
    rc = launch_score<MODE_DENSE, TF32>(a, ld_a, b, ld_b * L.stride, nb, L.n_s, d, L.sample, S, st, false);
    if (rc) return rc;
    const size_t ksmem = (size_t)L.n_s * sizeof(uint32_t);
    if (ksmem > 32 * 1024)
      FRX_CUDA(cudaFuncSetAttribute(row_kth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksmem));
    row_kth_kernel<<<nb, 256, ksmem, st>>>(sdense, L.n_s, (int)L.n_s, k, P.row_thr);
    FRX_LAUNCH_CHECK();
  } else if (L.has_sample) {
    const int64_t n_s = L.n_s, stride = L.stride;
    const Plan& ps = L.sample;
    ScoreParams S = P;
    S.labels = nullptr;
    S.pos_score = nullptr;
    S.index_base = 0;
    S.cap = ps.cap;
    S.keep_limit = ps.keep_limit;
    rc = launch_score<MODE_TOPK, TF32>(a, ld_a, b, ld_b * stride, nb, n_s, d, ps, S, st, false);
    if (rc) return rc;
    rc = run_merge(ps);
    if (rc) return rc;
    seed_threshold_kernel<<<(nb + 255) / 256, 256, 0, st>>>(topk_scores, nb, k, P.row_thr);
    FRX_LAUNCH_CHECK();
  }
  // The seeded thresholds are the origin of the per-row candidate histograms through which the CTAs of a row share
  // how many scores above each level have been seen so far (see hist_edge): the append threshold then tracks the
  // k-th best of everything processed by ANY CTA, not only of the CTA's own slice.
  if (L.has_sample) {
    FRX_CUDA(cudaMemcpyAsync(row_base, P.row_thr, L.thr_bytes, cudaMemcpyDeviceToDevice, st));
    FRX_CUDA(cudaMemsetAsync(row_hist, 0, L.hist_bytes, st));
    P.row_base = row_base;
    P.row_hist = row_hist;
    static const int refine_env = getenv("FRX_REFINE_EVERY") ? atoi(getenv("FRX_REFINE_EVERY")) : 0;   // tuning knob
    P.refine_every = refine_env >= 2 ? refine_env : REFINE_EVERY;
  }
  if (dense_out) {   // the main pass writes the scores from the same accumulators (no second contraction)
    FRX_CHECK_ARG(ld_dense >= n_posts, "frx_score_topk: ld_dense %lld below n_posts", (long long)ld_dense);
    P.dense = dense_out;
    P.ld_dense = ld_dense;
  }
  rc = launch_score<MODE_TOPK, TF32>(a, ld_a, b, ld_b, nb, n_posts, d, plan, P, st);
  if (rc) return rc;
  return run_merge(plan);
}

// out[m, n] = scale * sum_k A[m, k] * B[n, k] on the tf32 tensor-core path (used by the 3xTF32 brand embedding).
int dense_tf32_scaled(const float* a, int64_t ld_a, const float* b, int64_t ld_b, int m, int64_t n, int k, float* out,
                      int64_t ld_out, float scale, void* stream, void* ksplit_ws, size_t ksplit_ws_bytes) {
  return score_dense_impl<true>(a, ld_a, b, ld_b, m, n, k, out, ld_out, stream, scale, ksplit_ws, ksplit_ws_bytes);
}

}  // namespace frx

extern "C" {

size_t frx_score_topk_workspace_bytes(int nb, int64_t n_posts, int d, int k) {
  (void)d;
  if (nb <= 0 || n_posts <= 0 || k <= 0 || k > 1024) return 0;
  const size_t a = frx::make_topk_layout(nb, n_posts, k, false).total(), b = frx::make_topk_layout(nb, n_posts, k, true).total();
  return a > b ? a : b;     // one query serves the bf16 and the tf32 entry points (their kernel variants may differ)
}

int frx_score_topk(const uint16_t* brand_bf16, int64_t ld_a, const uint16_t* post_bf16, int64_t ld_b, int nb,
                   int64_t n_posts, int d, int k, const int32_t* labels, int64_t index_base, float* topk_scores,
                   int32_t* topk_index, float* pos_score, float* dense_out, int64_t ld_dense, void* workspace,
                   size_t workspace_bytes, void* stream) {
  return frx::score_topk_impl<false>(brand_bf16, ld_a, post_bf16, ld_b, nb, n_posts, d, k, labels, index_base, topk_scores,
                                     topk_index, pos_score, dense_out, ld_dense, workspace, workspace_bytes, stream);
}
int frx_score_topk_tf32(const float* brand_f32, int64_t ld_a, const float* post_f32, int64_t ld_b, int nb,
                        int64_t n_posts, int d, int k, const int32_t* labels, int64_t index_base, float* topk_scores,
                        int32_t* topk_index, float* pos_score, float* dense_out, int64_t ld_dense, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return frx::score_topk_impl<true>(brand_f32, ld_a, post_f32, ld_b, nb, n_posts, d, k, labels, index_base, topk_scores,
                                    topk_index, pos_score, dense_out, ld_dense, workspace, workspace_bytes, stream);
}

int frx_score_dense(const uint16_t* brand_bf16, int64_t ld_a, const uint16_t* post_bf16, int64_t ld_b, int nb,
                    int64_t n_posts, int d, float* dense_out, int64_t ld_dense, void* stream) {
  return frx::score_dense_impl<false>(brand_bf16, ld_a, post_bf16, ld_b, nb, n_posts, d, dense_out, ld_dense, stream);
}
int frx_score_dense_tf32(const float* brand_f32, int64_t ld_a, const float* post_f32, int64_t ld_b, int nb,
                         int64_t n_posts, int d, float* dense_out, int64_t ld_dense, void* stream) {
  return frx::score_dense_impl<true>(brand_f32, ld_a, post_f32, ld_b, nb, n_posts, d, dense_out, ld_dense, stream);
}

int frx_score_count(const uint16_t* brand_bf16, int64_t ld_a, const uint16_t* post_bf16, int64_t ld_b, int nb,
                    int64_t n_posts, int d, int64_t index_base, const float* thr_score, const int32_t* thr_index,
                    unsigned long long* count_out, void* stream) {
  return frx::score_count_impl<false>(brand_bf16, ld_a, post_bf16, ld_b, nb, n_posts, d, index_base, thr_score, thr_index,
                                      count_out, stream);
}
int frx_score_count_tf32(const float* brand_f32, int64_t ld_a, const float* post_f32, int64_t ld_b, int nb,
                         int64_t n_posts, int d, int64_t index_base, const float* thr_score, const int32_t* thr_index,
                         unsigned long long* count_out, void* stream) {
  return frx::score_count_impl<true>(brand_f32, ld_a, post_f32, ld_b, nb, n_posts, d, index_base, thr_score, thr_index,
                                     count_out, stream);
}

#ifdef FRX_TRACE
int frx_debug_set_trace(long long* device_buf) { frx::g_trace = device_buf; return 0; }
int frx_debug_set_cta_trace(long long* device_buf) { frx::g_cta_trace = device_buf; return 0; }
#endif

int frx_set_cta_pairs(int on) {
  const int prev = frx::use_pair() ? 1 : 0;
  frx::g_pair_mode = on < 0 ? -1 : (on != 0 ? 1 : 0);
  return prev;
}

int frx_set_cluster(int cta_count) {
  const int prev = frx::use_cluster();
  frx::g_cluster_mode = cta_count < 0 ? -1 : ((cta_count == 2 || cta_count == 4 || cta_count == 8) ? cta_count : 0);
  return prev;
}

int frx_probe_enable(int on) {
  frx::g_probe.on = on != 0;
  frx::g_probe.n = 0;
  return FRX_OK;
}

int frx_probe_read(float* host_ms_out, int max) {
  using namespace frx;
  int n = g_probe.n < max ? g_probe.n : max;
  for (int i = 0; i < n; ++i) {
    FRX_CUDA(cudaEventSynchronize(g_probe.end[i]));
    FRX_CUDA(cudaEventElapsedTime(host_ms_out + i, g_probe.beg[i], g_probe.end[i]));
  }
  g_probe.n = 0;
  return n;
}

int frx_topk_merge_strided(const float* in_scores, const int32_t* in_index, int g, int nb, int k_in, int64_t shard_stride,
                           float* out_scores, int32_t* out_index, int k_out, void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(in_scores && in_index && out_scores && out_index, "frx_topk_merge: NULL pointer");
  FRX_CHECK_ARG(g >= 1 && nb >= 1 && k_in >= 1 && k_out >= 1, "frx_topk_merge: bad sizes");
  FRX_CHECK_ARG(shard_stride >= (int64_t)nb * k_in, "frx_topk_merge: shard stride %lld below nb * k_in", (long long)shard_stride);
  FRX_CHECK_ARG((long)g * k_in <= MAX_MERGE_KEYS, "frx_topk_merge: g*k_in = %ld exceeds %d", (long)g * k_in, MAX_MERGE_KEYS);
  const size_t msmem = ((size_t)g * k_in + (size_t)k_out) * sizeof(unsigned long long);
  FRX_CHECK_ARG(msmem <= 200 * 1024, "frx_topk_merge: g*k_in + k_out = %zu keys exceed shared memory", msmem / 8);
  if (msmem > 32 * 1024)
    FRX_CUDA(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
  merge_lists_kernel<<<nb, 256, msmem, (cudaStream_t)stream>>>(in_scores, in_index, g, nb, k_in, shard_stride, out_scores,
                                                               out_index, k_out);
  FRX_LAUNCH_CHECK();
  return FRX_OK;
}

int frx_topk_merge(const float* in_scores, const int32_t* in_index, int g, int nb, int k_in, float* out_scores,
                   int32_t* out_index, int k_out, void* stream) {
  return frx_topk_merge_strided(in_scores, in_index, g, nb, k_in, (int64_t)nb * k_in, out_scores, out_index, k_out, stream);
}

}  // extern "C"
