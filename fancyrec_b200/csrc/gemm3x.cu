// fp32-grade GEMM on the tf32 tensor cores with the 3xTF32 operand split done INSIDE the kernel.
//
//   C[m, n] = act( alpha * (sum_k A(m, k) * B(n, k)) * col_scale[n] + col_shift[n] )
//
// A and B are plain fp32 arrays in global memory, each either K-major (X[row * ld + k]) or MN-major (X[k * ld + row],
// i.e. the transposed view) -- so S = post . brand^T, dPost = dS . brand, dBrand = dS^T . post (loss.py:87-143) and the
// Linear layers y = x . W^T + b of the encoders (model.py:59-83, 463-491) all run without a transposed or split copy of
// an operand ever touching HBM.  Round 1 ran these as  split kernel x 2 -> tf32 GEMM over K-concatenated [hi|lo|hi]
// operands (3x the operand bytes) -> k-split reduction: 4 launches and ~7x the operand traffic per product.
//
// One CTA = one 128 x 128 output tile (x one K range when K is split), 384 threads, warp-specialised:
//   warp 0 (one lane)   TMA producer: raw fp32 tiles {A 128 x 32, B 128 x 32} into a 3-stage ring, always in a layout the
//                       MMA can read: K-major sources as [128 rows][32 k] with SWIZZLE_128B; MN-major sources (row count a
//                       multiple of 32) as [4 groups of 32 rows][32 k][32 rows] with SWIZZLE_128B_ATOM_32B through a 3-D
//                       tensor map {row in group, k, group} -- the tensor core's MN-major tf32 operand layout
//                       (instruction descriptor bits 15 / 16), so nothing is transposed.  MN-major sources with a ragged row
//                       count fall back to plain [32 k][128 rows] tiles that the converters transpose;
//   warps 4-11          converters: x -> hi = tf32(x), lo = x - hi (Dekker split), written as FOUR tiles (A_hi, A_lo,
//                       B_hi, B_lo) at the offsets of the raw tile (elementwise, 16-byte accesses, conflict-free;
//                       fallback: lane = row, four k-consecutive scalar reads -> one swizzled 16-byte write), 2-stage
//                       ring, fence.proxy.async;
//   warp 1 (one lane)   tcgen05.mma kind::tf32, M = N = 128, K = 8: per k-block 3 products x 4 instructions
//                       (hi.hi + lo.hi + hi.lo), fp32 accumulation in 128 TMEM columns;
//   warps 4-11 again    epilogue: tcgen05.ld -> scale / shift / ReLU -> C, or the raw partial tile of a K split.
// Up to two independent problems share one launch (the two gradient GEMMs; the raw and the normalised tile).
#include <cstdlib>

#include "common.cuh"
#include "sm100.cuh"

namespace frx {
using namespace sm100;
namespace g3 {

constexpr int TM = 128, TN = 128, KB = 32;                 // output tile; k-block = 32 fp32 = one 128-byte swizzle row
constexpr int TILE_BYTES = TM * KB * 4;                    // 16 KB per operand tile (TM == TN)
constexpr int NRAW = 3, NCONV = 2;
constexpr int CONV_WARP0 = 4, NUM_CONV_WARPS = 8, NUM_THREADS = (CONV_WARP0 + NUM_CONV_WARPS) * 32;
constexpr int CONV_THREADS = NUM_CONV_WARPS * 32;
constexpr int TMEM_COLS = 128;

struct Tail {
  uint64_t raw_full[NRAW], raw_empty[NRAW], conv_full[NCONV], conv_empty[NCONV], acc_full;
  uint32_t tmem_base;
};
constexpr size_t SMEM_BYTES = 1024 + (size_t)NRAW * 2 * TILE_BYTES + (size_t)NCONV * 4 * TILE_BYTES + sizeof(Tail);

struct Problem {
  float* c; int64_t ldc;
  float* partial;                  // [ksplit][m][n] raw accumulators when ksplit > 1 (or when the caller wants them)
  const float* col_scale; const float* col_shift;
  float alpha;
  int m, n, k, a_mn, b_mn, relu;        // a_mn / b_mn: 0 = K-major, 1 = MN-major transposed by the converters, 2 = MN-major read in place
  int tiles_m, tiles_n, ksplit, cta0;   // tile-per-CTA mapping: this problem's CTAs are [cta0, cta0 + tiles_m * tiles_n * ksplit)
  int nkb, unit0;                       // k-blocks per tile; stream mapping: this problem's first (tile, k-block) unit
};
// Two ways of handing the (tile, k-block) units to CTAs:
//   stream == 0: one CTA = one tile x one of `ksplit` equal K ranges (grid = tiles * ksplit);
//   stream == 1: the units of all tiles, tile-major, are cut into gridDim.x EQUAL contiguous ranges (one CTA per SM), so
//                160 tiles on 148 SMs cost 160 / 148 of a tile each instead of two waves.  A CTA finishes the tiles that
//                lie wholly inside its range; the at most two tiles it shares with its neighbours go, as raw
//                accumulators, to its two slots in `slots` and `fixup_kernel` adds a tile's pieces in CTA order.
struct Params { Problem p[2]; int count, stream, total_units; float* slots; };

// A CTA's share of the unit stream, handed out as pieces = runs of k-blocks of ONE tile (the inner loops of all three
// roles run over one piece with everything else loop-invariant).
struct Walker { int p, tile, kb, left; };
struct Piece { int p, tile, kb0, kb1; };
__device__ __forceinline__ int64_t range_begin(int cta, int ncta, int total) { return (int64_t)cta * total / ncta; }
__device__ __forceinline__ Walker walker_init(const Params& P, int cta, int ncta) {
  Walker w;
  const int u0 = (int)range_begin(cta, ncta, P.total_units), u1 = (int)range_begin(cta + 1, ncta, P.total_units);
  w.left = u1 - u0;
  w.p = (P.count > 1 && u0 >= P.p[1].unit0) ? 1 : 0;
  const int local = u0 - P.p[w.p].unit0;
  w.tile = local / P.p[w.p].nkb;
  w.kb = local - w.tile * P.p[w.p].nkb;
  return w;
}
__device__ __forceinline__ bool next_piece(const Params& P, Walker& w, Piece& pc) {
  if (w.left <= 0) return false;
  const int nkb = P.p[w.p].nkb;
  const int n = nkb - w.kb < w.left ? nkb - w.kb : w.left;
  pc.p = w.p; pc.tile = w.tile; pc.kb0 = w.kb; pc.kb1 = w.kb + n;
  w.left -= n;
  w.kb += n;
  if (w.kb == nkb) {
    w.kb = 0;
    if (++w.tile == P.p[w.p].tiles_m * P.p[w.p].tiles_n) { w.tile = 0; ++w.p; }
  }
  return true;
}

// x = hi + lo with hi exactly representable in tf32 (11 significant bits, round to nearest): Dekker's split with the
// constant 2^13 + 1 -- three full-rate fp32 operations (cvt.rna.tf32.f32 runs on the quarter-rate conversion pipe and
// made the converters, not the tensor core, the bottleneck: 1.1 us per k-block).  lo = x - hi is exact in fp32 and has
// at most 13 significant bits; the tensor core reads its leading 11, so x is represented to 2^-22 relative.
__device__ __forceinline__ void split1(float x, float& h, float& l) {
  const float p = __fmul_rn(x, 8193.0f);          // _rn intrinsics: never contracted into an FMA (which would skip
  h = __fsub_rn(p, __fsub_rn(p, x));              // the rounding of p that the split relies on)
  l = __fsub_rn(x, h);
}
__device__ __forceinline__ void split4(const float4& v, float4& h, float4& l) {
  split1(v.x, h.x, l.x);
  split1(v.y, h.y, l.y);
  split1(v.z, h.z, l.z);
  split1(v.w, h.w, l.w);
}
// raw tile already in a layout the MMA reads (K-major or MN-major in place): the split is elementwise, offsets carry over
__device__ __forceinline__ void convert_kmajor(const uint8_t* raw, uint8_t* hi, uint8_t* lo, int t) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int off = (t + i * CONV_THREADS) * 16;
    float4 h, l;
    split4(*reinterpret_cast<const float4*>(raw + off), h, l);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}
// raw tile = [32 k][128 rows] fp32 (the transposed view of the operand): lane = row, four k-consecutive scalar reads
// (conflict-free: consecutive lanes, consecutive words) -> the 16-byte chunk kc of row r at its swizzled position
// r/8 * 1024 + r%8 * 128 + ((kc ^ r%8) << 4)  (8 consecutive lanes cover 8 distinct 16-byte bank groups)
__device__ __forceinline__ void convert_mnmajor(const uint8_t* raw8, uint8_t* hi, uint8_t* lo, int t) {
  const float* raw = reinterpret_cast<const float*>(raw8);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = t + i * CONV_THREADS;
    const int r = c & (TM - 1), kc = c >> 7;
    float4 v, h, l;
    v.x = raw[(4 * kc + 0) * TM + r];
    v.y = raw[(4 * kc + 1) * TM + r];
    v.z = raw[(4 * kc + 2) * TM + r];
    v.w = raw[(4 * kc + 3) * TM + r];
    split4(v, h, l);
    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((kc ^ (r & 7)) << 4);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

// STREAM = false: one CTA = one piece whose problem / tile / K range follow from blockIdx alone -- everything the MMA
// issuer needs stays in uniform registers.  STREAM = true: the CTA walks its unit range piece by piece (descriptors and
// layout flags then live in vector registers: one ELECT / R2UR round trip per MMA, ~5 % per k-block).
template <bool STREAM>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_b0,
              const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
              const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  uint8_t* raw_tiles = smem;                                        // [NRAW][A | B]
  uint8_t* conv_tiles = smem + (size_t)NRAW * 2 * TILE_BYTES;       // [NCONV][A_hi | A_lo | B_hi | B_lo]
  Tail* tail = reinterpret_cast<Tail*>(conv_tiles + (size_t)NCONV * 4 * TILE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int cta = (int)blockIdx.x, ncta = (int)gridDim.x;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a0);
    prefetch_tmap(&map_b0);
    if (P.count > 1) { prefetch_tmap(&map_a1); prefetch_tmap(&map_b1); }
    for (int s = 0; s < NRAW; ++s) { mbar_init(smem_u32(&tail->raw_full[s]), 1); mbar_init(smem_u32(&tail->raw_empty[s]), NUM_CONV_WARPS); }
    for (int s = 0; s < NCONV; ++s) { mbar_init(smem_u32(&tail->conv_full[s]), NUM_CONV_WARPS); mbar_init(smem_u32(&tail->conv_empty[s]), 1); }
    mbar_init(smem_u32(&tail->acc_full), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(&tail->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  Walker w{};
  if constexpr (STREAM) w = walker_init(P, cta, ncta);
  Piece pc;
  if constexpr (!STREAM) {                                     // the single piece, in the terms of the launch (uniform)
    pc.p = (P.count > 1 && cta >= P.p[1].cta0) ? 1 : 0;
    const int local = cta - P.p[pc.p].cta0, ks = local % P.p[pc.p].ksplit;
    pc.tile = local / P.p[pc.p].ksplit;
    pc.kb0 = P.p[pc.p].nkb * ks / P.p[pc.p].ksplit;
    pc.kb1 = P.p[pc.p].nkb * (ks + 1) / P.p[pc.p].ksplit;
  }
  auto first_piece = [&]() { if constexpr (STREAM) return next_piece(P, w, pc); else return true; };
  auto another_piece = [&]() { if constexpr (STREAM) return next_piece(P, w, pc); else return false; };
  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (bool have = first_piece(); have; have = another_piece()) {
        const Problem& Q = P.p[pc.p];
        const CUtensorMap* map_a = pc.p ? &map_a1 : &map_a0;
        const CUtensorMap* map_b = pc.p ? &map_b1 : &map_b0;
        const int tn = pc.tile % Q.tiles_n, tm = pc.tile / Q.tiles_n;
        const int a_mn = Q.a_mn, b_mn = Q.b_mn;
        for (int kb = pc.kb0; kb < pc.kb1; ++kb) {
          mbar_wait(smem_u32(&tail->raw_empty[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&tail->raw_full[stage]);
          mbar_arrive_expect_tx(fb, 2 * TILE_BYTES);
          const uint32_t dst_a = base + stage * 2 * TILE_BYTES, dst_b = dst_a + TILE_BYTES;
          if (a_mn == 2) tma_load_3d(dst_a, map_a, fb, 0, kb * KB, tm * (TM / 32));
          else if (a_mn) tma_load_2d(dst_a, map_a, fb, tm * TM, kb * KB);
          else tma_load_2d(dst_a, map_a, fb, kb * KB, tm * TM);
          if (b_mn == 2) tma_load_3d(dst_b, map_b, fb, 0, kb * KB, tn * (TN / 32));
          else if (b_mn) tma_load_2d(dst_b, map_b, fb, tn * TN, kb * KB);
          else tma_load_2d(dst_b, map_b, fb, kb * KB, tn * TN);
          if (++stage == NRAW) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t GROUP_BYTES = KB * 128;                   // one 32-row MN group: [32 k][128 bytes]
      int cs = 0; uint32_t cphase = 0;
      for (bool have = first_piece(); have; have = another_piece()) {
        const bool a_dir = P.p[pc.p].a_mn == 2, b_dir = P.p[pc.p].b_mn == 2;
        const uint32_t idesc = make_idesc_tf32(TM, TN) | (a_dir ? 1u << 15 : 0u) | (b_dir ? 1u << 16 : 0u);
        // one K = 8 step: 32 bytes along a K-major row, 8 k rows (1024 bytes) of an MN-major tile (>> 4 in the descriptor)
        const uint64_t sa = a_dir ? 64 : 2, sb = b_dir ? 64 : 2;
        const int n = pc.kb1 - pc.kb0;
        for (int i = 0; i < n; ++i) {
          mbar_wait(smem_u32(&tail->conv_full[cs]), cphase);
          tc_fence_after();
          const uint32_t t0 = base + NRAW * 2 * TILE_BYTES + cs * 4 * TILE_BYTES;
          auto desc = [&](uint32_t addr, bool dir) { return dir ? make_sw128_mnmajor_desc(addr, GROUP_BYTES) : make_sw128_kmajor_desc(addr); };
          const uint64_t a_hi = desc(t0, a_dir), a_lo = desc(t0 + TILE_BYTES, a_dir);
          const uint64_t b_hi = desc(t0 + 2 * TILE_BYTES, b_dir), b_lo = desc(t0 + 3 * TILE_BYTES, b_dir);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32_ss(tmem_base, a_lo + sa * k, b_hi + sb * k, idesc, (i > 0 || k > 0) ? 1u : 0u);   // small terms first
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32_ss(tmem_base, a_hi + sa * k, b_lo + sb * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32_ss(tmem_base, a_hi + sa * k, b_hi + sb * k, idesc, 1u);
          umma_commit(smem_u32(&tail->conv_empty[cs]));            // the four split tiles are free when these retire
          if (i == n - 1) umma_commit(smem_u32(&tail->acc_full));
          if (++cs == NCONV) { cs = 0; cphase ^= 1; }
        }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // =========================== converters, and the epilogue of every piece ===========================
    const int t = threadIdx.x - CONV_WARP0 * 32;
    const int q = warp & 3, h = (warp - CONV_WARP0) >> 2;          // epilogue: TMEM lane quarter, column half
    int rs = 0; uint32_t rphase = 0;
    int cs = 0; uint32_t cphase = 0;
    uint32_t aphase = 0;
    bool cta_first = true;
    for (bool have = first_piece(); have; have = another_piece()) {
      const Problem& Q = P.p[pc.p];
      const bool a_tr = Q.a_mn == 1, b_tr = Q.b_mn == 1;
      for (int kb = pc.kb0; kb < pc.kb1; ++kb) {
        mbar_wait(smem_u32(&tail->raw_full[rs]), rphase);
        mbar_wait(smem_u32(&tail->conv_empty[cs]), cphase ^ 1);
        const uint8_t* ra = raw_tiles + (size_t)rs * 2 * TILE_BYTES;
        const uint8_t* rb = ra + TILE_BYTES;
        uint8_t* c0 = conv_tiles + (size_t)cs * 4 * TILE_BYTES;
        if (a_tr) convert_mnmajor(ra, c0, c0 + TILE_BYTES, t); else convert_kmajor(ra, c0, c0 + TILE_BYTES, t);
        if (b_tr) convert_mnmajor(rb, c0 + 2 * TILE_BYTES, c0 + 3 * TILE_BYTES, t);
        else convert_kmajor(rb, c0 + 2 * TILE_BYTES, c0 + 3 * TILE_BYTES, t);
        fence_proxy_async();                                       // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&tail->conv_full[cs]));
          mbar_arrive(smem_u32(&tail->raw_empty[rs]));
        }
        if (++rs == NRAW) { rs = 0; rphase ^= 1; }
        if (++cs == NCONV) { cs = 0; cphase ^= 1; }
      }
      // ---- epilogue of the piece: thread = TMEM lane = one output row; two warps per lane quarter split the 128 columns.
      // The MMA issuer cannot start the next piece before every converter warp has delivered that piece's first
      // k-block, i.e. after all of them have left this epilogue: the accumulator needs no barrier of its own.
      const int tn = pc.tile % Q.tiles_n, tm = pc.tile / Q.tiles_n;
      const bool whole = pc.kb0 == 0 && pc.kb1 == Q.nkb;
      const int row = tm * TM + q * 32 + lane;
      float* piece = nullptr;                                      // raw accumulators of a shared tile (stream mapping)
      if (STREAM && !whole) piece = P.slots + ((size_t)cta * 2 + (cta_first ? 0 : 1)) * (TM * TN) + (size_t)(q * 32 + lane) * TN;
      cta_first = false;
      mbar_wait(smem_u32(&tail->acc_full), aphase);
      aphase ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        const int col0 = tn * TN + h * 64 + cc * 32;
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64 + cc * 32), v);
        tmem_ld_wait();
        if (piece != nullptr) {
          float* dst = piece + h * 64 + cc * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                              __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          continue;
        }
        if (row >= Q.m || col0 >= Q.n) continue;
        const int nvalid = Q.n - col0 >= 32 ? 32 : Q.n - col0;
        if (Q.partial != nullptr) {
          const int ks = (cta - Q.cta0) % Q.ksplit;
          float* dst = Q.partial + ((int64_t)ks * Q.m + row) * Q.n + col0;
          if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < nvalid) dst[i] = __uint_as_float(v[i]);
          }
        } else {
          float* dst = Q.c + (int64_t)row * Q.ldc + col0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < nvalid) {
              float o = __uint_as_float(v[i]) * Q.alpha;
              if (Q.col_scale) o *= __ldg(Q.col_scale + col0 + i);
              if (Q.col_shift) o += __ldg(Q.col_shift + col0 + i);
              if (Q.relu) o = fmaxf(o, 0.f);
              v[i] = __float_as_uint(o);
            }
          }
          if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < nvalid) dst[i] = __uint_as_float(v[i]);
          }
        }
      }
      tc_fence_before();                                           // TMEM reads ordered before the arrivals that let the next piece start
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// C = act(alpha * sum_ks partial[ks] * col_scale + col_shift): the K-split reduction, summed in k-split order
__global__ void __launch_bounds__(256) reduce_partial_kernel(const float* __restrict__ partial, int ksplit, int m, int n,
                                                             float* __restrict__ c, int64_t ldc, float alpha,
                                                             const float* __restrict__ col_scale,
                                                             const float* __restrict__ col_shift, int relu) {
  const int64_t total = (int64_t)m * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / n), col = (int)(i - (int64_t)r * n);
    float acc = 0.f;
    for (int ks = 0; ks < ksplit; ++ks) acc += partial[(int64_t)ks * total + i];
    float o = acc * alpha;
    if (col_scale) o *= col_scale[col];
    if (col_shift) o += col_shift[col];
    if (relu) o = fmaxf(o, 0.f);
    c[(int64_t)r * ldc + col] = o;
  }
}

// Stream mapping: FIX_SPLIT blocks per output tile (32 rows each).  A tile whose units all fell into one CTA's range was
// finished by that CTA; otherwise its pieces lie in the slots of the CTAs first .. last that share it (slot 0 = the piece a
// CTA's range starts with, slot 1 = the piece it ends with) and are added here in CTA order -- a fixed order, so results
// are reproducible.
constexpr int FIX_SPLIT = 4, FIX_MAX_PIECES = 160;
__global__ void __launch_bounds__(256) fixup_kernel(const __grid_constant__ Params P, int ncta) {
  __shared__ int s_slot[FIX_MAX_PIECES];
  int tile = (int)blockIdx.x / FIX_SPLIT, p = 0;
  const int part = (int)blockIdx.x % FIX_SPLIT;
  if (P.count > 1 && tile >= P.p[0].tiles_m * P.p[0].tiles_n) { tile -= P.p[0].tiles_m * P.p[0].tiles_n; p = 1; }
  const Problem& Q = P.p[p];
  const int64_t u0 = Q.unit0 + (int64_t)tile * Q.nkb, u1 = u0 + Q.nkb;           // this tile's units
  const int first = (int)(((u0 + 1) * ncta - 1) / P.total_units);               // CTA whose range holds unit u0
  const int last = (int)((u1 * ncta - 1) / P.total_units);                      // ... unit u1 - 1
  if (first == last) return;
  const int np = last - first + 1;
  for (int j = threadIdx.x; j < np; j += 256) s_slot[j] = (first + j) * 2 + (range_begin(first + j, ncta, P.total_units) >= u0 ? 0 : 1);
  __syncthreads();
  constexpr int PART = TM * TN / FIX_SPLIT;                                      // floats per block: 32 rows
  constexpr int PER = PART / 4 / 256;                                            // float4 per thread
  float4 acc[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < np; ++j) {
    const float4* src = reinterpret_cast<const float4*>(P.slots + (size_t)s_slot[j] * (TM * TN) + (size_t)part * PART);
    float4 v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) v[u] = __ldcg(src + threadIdx.x + u * 256);
#pragma unroll
    for (int u = 0; u < PER; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
  }
  const int tn = tile % Q.tiles_n, tm = tile / Q.tiles_n;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int e = part * PART + (threadIdx.x + u * 256) * 4;
    const int r = e / TN, col = e - r * TN;
    const int row = tm * TM + r, col0 = tn * TN + col;
    if (row >= Q.m || col0 >= Q.n) continue;
    float o[4] = {acc[u].x, acc[u].y, acc[u].z, acc[u].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i] *= Q.alpha;
      if (col0 + i < Q.n) {
        if (Q.col_scale) o[i] *= __ldg(Q.col_scale + col0 + i);
        if (Q.col_shift) o[i] += __ldg(Q.col_shift + col0 + i);
      }
      if (Q.relu) o[i] = fmaxf(o[i], 0.f);
    }
    float* dst = Q.c + (int64_t)row * Q.ldc + col0;
    if (col0 + 4 <= Q.n && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      for (int i = 0; i < 4 && col0 + i < Q.n; ++i) dst[i] = o[i];
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
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

// operand X(row, k): K-major (0) = x[row * ld + k] -> dims {k, rows}, box {32, 128}, SWIZZLE_128B;
//   MN-major transposed by the converters (1) = x[k * ld + row] -> dims {rows, k}, box {128, 32}, no swizzle;
//   MN-major read in place (2, rows % 32 == 0) -> dims {32 rows of a group, k, rows / 32 groups} with byte strides {ld * 4, 128},
//   box {32, 32, 4}, SWIZZLE_128B_ATOM_32B: shared memory holds [group][k][32 rows].  OOB reads are zero.
static int make_map(CUtensorMap* map, const float* x, int64_t ld, int rows, int k, int mode) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return FRX_E_DEVICE; }
  cuuint64_t dims[3], strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  int rank = 2;
  if (mode == 2) { rank = 3; dims[0] = 32; dims[1] = (cuuint64_t)k; dims[2] = (cuuint64_t)(rows / 32); box[0] = 32; box[1] = KB; box[2] = TM / 32; }
  else if (mode == 1) { dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)k; box[0] = TM; box[1] = KB; }
  else { dims[0] = (cuuint64_t)k; dims[1] = (cuuint64_t)rows; box[0] = KB; box[1] = TM; }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(x), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mode == 1 ? CU_TENSOR_MAP_SWIZZLE_NONE : mode == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (gemm3x) failed with CUresult %d", (int)r); return FRX_E_CUDA; }
  return FRX_OK;
}

}  // namespace g3

bool gemm3x_supported(const Gemm3xDesc& d) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return d.m > 0 && d.n > 0 && d.k > 0 && d.lda % 4 == 0 && d.ldb % 4 == 0 && al(d.a) && al(d.b) &&
         d.lda >= (d.a_mn ? d.m : d.k) && d.ldb >= (d.b_mn ? d.n : d.k);
}

int gemm3x_plan_ksplit(const Gemm3xDesc* d, int count) {
  long tiles = 0;
  int min_kb = 1 << 30;
  for (int i = 0; i < count; ++i) {
    tiles += (long)((d[i].m + g3::TM - 1) / g3::TM) * ((d[i].n + g3::TN - 1) / g3::TN);
    const int kb = (d[i].k + g3::KB - 1) / g3::KB;
    min_kb = kb < min_kb ? kb : min_kb;
  }
  const int sms = num_sms();
  int ks = 1;
  if (tiles * 2 <= sms) {
    ks = (int)(sms / tiles);
    if (ks > min_kb / 4) ks = min_kb / 4;      // at least 4 k-blocks per CTA
    if (ks < 1) ks = 1;
  }
  return ks;
}

size_t gemm3x_partial_floats(const Gemm3xDesc* d, int count, int ksplit) {
  size_t total = 0;
  for (int i = 0; i < count; ++i) total += (size_t)ksplit * (size_t)d[i].m * (size_t)d[i].n;
  return total;
}

// ---- stream mapping: when is it worth it, and how many CTAs
namespace g3 {
struct StreamPlan { int ctas; long units, tiles; };
static StreamPlan stream_plan(const Gemm3xDesc* d, int count) {
  StreamPlan sp{0, 0, 0};
  for (int i = 0; i < count; ++i) {
    const long t = (long)((d[i].m + TM - 1) / TM) * ((d[i].n + TN - 1) / TN);
    sp.tiles += t;
    sp.units += t * ((d[i].k + KB - 1) / KB);
  }
  const long g = sp.units / 8;                       // at least 8 k-blocks per CTA
  sp.ctas = (int)(g < num_sms() ? g : num_sms());
  return sp;
}
}  // namespace g3

size_t gemm3x_stream_floats(const Gemm3xDesc* d, int count) {
  const g3::StreamPlan sp = g3::stream_plan(d, count);
  return sp.ctas >= 2 ? (size_t)sp.ctas * 2 * g3::TM * g3::TN : 0;
}

// Estimated cost in k-block times (0.9 us each): a CTA's fixed cost ~ 5, a K-split reduction ~ 8, the stream mapping's two
// extra epilogues + fix-up ~ 14.  The stream mapping is taken when it is at least 10 % cheaper than whole tiles per CTA.
static bool stream_pays(const Gemm3xDesc* d, int count, int ksplit) {
  static const bool off = [] { const char* e = getenv("FRX_G3_STREAM"); return e && e[0] == '0'; }();
  if (off) return false;
  const g3::StreamPlan sp = g3::stream_plan(d, count);
  if (sp.ctas < 2) return false;
  const int sms = num_sms();
  long ctas = 0, longest = 0;
  for (int i = 0; i < count; ++i) {
    const long nkb = (d[i].k + g3::KB - 1) / g3::KB;
    const long ks = ksplit < nkb ? ksplit : nkb;
    ctas += (long)((d[i].m + g3::TM - 1) / g3::TM) * ((d[i].n + g3::TN - 1) / g3::TN) * ks;
    const long per = (nkb + ks - 1) / ks;
    longest = per > longest ? per : longest;
  }
  if (ctas % sms == 0) return false;
  const long whole = (ctas + sms - 1) / sms * (longest + 5) + (ksplit > 1 ? 8 : 0);
  const long stream = (sp.units + sp.ctas - 1) / sp.ctas + 14;
  return stream * 10 < whole * 9;
}

// Launch up to two problems in ONE grid.  Whole tiles per CTA: ksplit > 1 (or keep_partials) sends raw accumulators to
// `partial` ([ksplit][m][n] per problem, problem 1 behind problem 0); unless keep_partials, a reduction kernel applies the
// epilogue.  When the tiles do not fill whole waves of SMs and the workspace allows (gemm3x_stream_floats), the units are
// streamed over one CTA per SM instead (see Params) and `partial` holds the shared tiles' pieces.
int gemm3x_launch(cudaStream_t st, const Gemm3xDesc* d, int count, int ksplit, bool keep_partials, float* partial,
                  size_t partial_floats) {
  using namespace g3;
  if (count < 1 || count > 2) { set_error("gemm3x: %d problems (1 or 2 supported)", count); return FRX_E_ARG; }
  for (int i = 0; i < count; ++i)
    if (!gemm3x_supported(d[i])) { set_error("gemm3x: unsupported operand layout (16-byte aligned, ld %% 4 == 0 required)"); return FRX_E_ARG; }
  if (ksplit < 1) ksplit = 1;
  const bool stream = !keep_partials && partial != nullptr && partial_floats >= gemm3x_stream_floats(d, count) &&
                      stream_pays(d, count, ksplit);
  if (stream) ksplit = 1;
  const bool to_partial = ksplit > 1 || keep_partials;
  if (to_partial && (partial == nullptr || partial_floats < gemm3x_partial_floats(d, count, ksplit))) {
    set_error("gemm3x: partial workspace too small");
    return FRX_E_WORKSPACE;
  }
  CUtensorMap maps[4];
  Params P{};
  P.count = count;
  P.stream = stream ? 1 : 0;
  P.slots = partial;
  int cta = 0, unit = 0, tiles = 0;
  float* pp = partial;
  for (int i = 0; i < count; ++i) {
    // MN-major operands are read in place when their row count is a whole number of 32-row groups
    // (FRX_G3_TRANSPOSE=1 forces the converter-transposed tiles, for A / B comparisons)
    static const bool force_transpose = [] { const char* e = getenv("FRX_G3_TRANSPOSE"); return e && e[0] == '1'; }();
    const int a_mode = !d[i].a_mn ? 0 : (d[i].m % 32 == 0 && !force_transpose) ? 2 : 1;
    const int b_mode = !d[i].b_mn ? 0 : (d[i].n % 32 == 0 && !force_transpose) ? 2 : 1;
    int rc = make_map(&maps[2 * i], d[i].a, d[i].lda, d[i].m, d[i].k, a_mode);
    if (rc) return rc;
    rc = make_map(&maps[2 * i + 1], d[i].b, d[i].ldb, d[i].n, d[i].k, b_mode);
    if (rc) return rc;
    Problem& Q = P.p[i];
    Q.c = d[i].c; Q.ldc = d[i].ldc;
    Q.partial = to_partial ? pp : nullptr;
    Q.col_scale = d[i].col_scale; Q.col_shift = d[i].col_shift;
    Q.alpha = d[i].alpha; Q.relu = d[i].relu;
    Q.m = d[i].m; Q.n = d[i].n; Q.k = d[i].k; Q.a_mn = a_mode; Q.b_mn = b_mode;
    Q.tiles_m = (d[i].m + TM - 1) / TM; Q.tiles_n = (d[i].n + TN - 1) / TN;
    Q.nkb = (d[i].k + KB - 1) / KB;
    Q.ksplit = ksplit < Q.nkb ? ksplit : Q.nkb;
    if (to_partial && Q.ksplit != ksplit) { set_error("gemm3x: k = %d too short for a %d-way K split", d[i].k, ksplit); return FRX_E_ARG; }
    Q.cta0 = cta;
    Q.unit0 = unit;
    cta += Q.tiles_m * Q.tiles_n * Q.ksplit;
    unit += Q.tiles_m * Q.tiles_n * Q.nkb;
    tiles += Q.tiles_m * Q.tiles_n;
    pp += (size_t)ksplit * (size_t)d[i].m * (size_t)d[i].n;
  }
  P.total_units = unit;
  if (count == 1) { maps[2] = maps[0]; maps[3] = maps[1]; P.p[1] = P.p[0]; P.p[1].cta0 = cta; P.p[1].unit0 = unit; }
  static bool attr_set = false;
  if (!attr_set) {
    FRX_CUDA(cudaFuncSetAttribute(gemm3x_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    FRX_CUDA(cudaFuncSetAttribute(gemm3x_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  if (stream) {
    const int g = stream_plan(d, count).ctas;
    gemm3x_kernel<true><<<g, NUM_THREADS, SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], P);
    FRX_LAUNCH_CHECK();
    // a fix-up is needed only if some CTA range boundary falls inside a tile
    bool shared = false;
    for (int c = 1; c < g && !shared; ++c) {
      const long u = (long)c * unit / g;
      const Problem& Q = P.p[(count > 1 && u >= P.p[1].unit0) ? 1 : 0];
      shared = (u - Q.unit0) % Q.nkb != 0;
    }
    if (shared) {
      fixup_kernel<<<tiles * FIX_SPLIT, 256, 0, st>>>(P, g);
      FRX_LAUNCH_CHECK();
    }
    return FRX_OK;
  }
  gemm3x_kernel<false><<<cta, NUM_THREADS, SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], P);
  FRX_LAUNCH_CHECK();
  if (to_partial && !keep_partials) {
    const float* src = partial;
    for (int i = 0; i < count; ++i) {
      const int64_t total = (int64_t)d[i].m * d[i].n;
      int64_t blocks = (total + 255) / 256;
      if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
      reduce_partial_kernel<<<(int)blocks, 256, 0, st>>>(src, ksplit, d[i].m, d[i].n, d[i].c, d[i].ldc, d[i].alpha,
                                                         d[i].col_scale, d[i].col_shift, d[i].relu);
      FRX_LAUNCH_CHECK();
      src += (size_t)ksplit * (size_t)total;
    }
  }
  return FRX_OK;
}

}  // namespace frx

extern "C" {

size_t frx_linear_workspace_bytes(int m, int n, int k) {
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  frx::Gemm3xDesc d{};
  d.m = m; d.n = n; d.k = k;
  const int ks = frx::gemm3x_plan_ksplit(&d, 1);
  const size_t split = ks > 1 ? frx::gemm3x_partial_floats(&d, 1, ks) : 0, stream = frx::gemm3x_stream_floats(&d, 1);
  return (split > stream ? split : stream) * sizeof(float) + 256;
}

int frx_linear(const float* x, int64_t ld_x, const float* w, int64_t ld_w, const float* col_scale, const float* col_shift,
               int relu, int m, int n, int k, float* out, int64_t ld_out, void* workspace, size_t workspace_bytes,
               void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(x && w && out, "frx_linear: NULL pointer");
  FRX_CHECK_ARG(m > 0 && n > 0 && k > 0 && ld_out >= n, "frx_linear: bad sizes m=%d n=%d k=%d", m, n, k);
  int dev = 0;
  FRX_CUDA(cudaGetDevice(&dev));
  int rc = frx_device_check(dev);
  if (rc) return rc;
  Gemm3xDesc d{};
  d.a = x; d.lda = ld_x; d.a_mn = 0;
  d.b = w; d.ldb = ld_w; d.b_mn = 0;
  d.c = out; d.ldc = ld_out;
  d.m = m; d.n = n; d.k = k;
  d.alpha = 1.0f; d.col_scale = col_scale; d.col_shift = col_shift; d.relu = relu;
  FRX_CHECK_ARG(gemm3x_supported(d), "frx_linear: x and w must be 16-byte aligned with row pitches that are multiples of 4 floats");
  const int ks = gemm3x_plan_ksplit(&d, 1);
  if (workspace == nullptr || workspace_bytes < frx_linear_workspace_bytes(m, n, k)) {
    set_error("frx_linear: workspace %zu bytes, need %zu", workspace_bytes, frx_linear_workspace_bytes(m, n, k));
    return FRX_E_WORKSPACE;
  }
  return gemm3x_launch((cudaStream_t)stream, &d, 1, ks, false, reinterpret_cast<float*>(workspace),
                       workspace_bytes / sizeof(float));
}

int frx_matmul3x(const float* a, int64_t ld_a, int a_transposed, const float* b, int64_t ld_b, int b_transposed, int m,
                 int n, int k, float alpha, float* out, int64_t ld_out, void* workspace, size_t workspace_bytes,
                 void* stream) {
  using namespace frx;
  FRX_CHECK_ARG(a && b && out, "frx_matmul3x: NULL pointer");
  FRX_CHECK_ARG(m > 0 && n > 0 && k > 0 && ld_out >= n, "frx_matmul3x: bad sizes m=%d n=%d k=%d", m, n, k);
  int dev = 0;
  FRX_CUDA(cudaGetDevice(&dev));
  int rc = frx_device_check(dev);
  if (rc) return rc;
  Gemm3xDesc d{};
  d.a = a; d.lda = ld_a; d.a_mn = a_transposed ? 1 : 0;
  d.b = b; d.ldb = ld_b; d.b_mn = b_transposed ? 1 : 0;
  d.c = out; d.ldc = ld_out;
  d.m = m; d.n = n; d.k = k;
  d.alpha = alpha;
  FRX_CHECK_ARG(gemm3x_supported(d), "frx_matmul3x: operands must be 16-byte aligned with pitches that are multiples of 4 floats");
  const int ks = gemm3x_plan_ksplit(&d, 1);
  if (workspace == nullptr || workspace_bytes < frx_linear_workspace_bytes(m, n, k)) {
    set_error("frx_matmul3x: workspace %zu bytes, need %zu", workspace_bytes, frx_linear_workspace_bytes(m, n, k));
    return FRX_E_WORKSPACE;
  }
  return gemm3x_launch((cudaStream_t)stream, &d, 1, ks, false, reinterpret_cast<float*>(workspace),
                       workspace_bytes / sizeof(float));
}

}  // extern "C"
