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
//   warp 0 (one lane)   TMA producer: raw fp32 tiles {A 128 x 32, B 128 x 32} into a 3-stage ring.  K-major sources are
//                       loaded with SWIZZLE_128B (the layout the MMA reads), MN-major sources as plain [32 k][128 rows];
//   warps 4-11          converters: x -> hi = tf32(x), lo = x - hi (Dekker split), written as FOUR tiles (A_hi, A_lo, B_hi, B_lo) in
//                       the SWIZZLE_128B K-major layout (MN-major sources are transposed on the way: lane = row, four
//                       k-consecutive scalar reads -> one swizzled 16-byte write), 2-stage ring, fence.proxy.async;
//   warp 1 (one lane)   tcgen05.mma kind::tf32, M = N = 128, K = 8: per k-block 3 products x 4 instructions
//                       (hi.hi + lo.hi + hi.lo), fp32 accumulation in 128 TMEM columns;
//   warps 4-11 again    epilogue: tcgen05.ld -> scale / shift / ReLU -> C, or the raw partial tile of a K split.
// Up to two independent problems share one launch (the two gradient GEMMs; the raw and the normalised tile).
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
  int m, n, k, a_mn, b_mn, relu;
  int tiles_m, tiles_n, ksplit, cta0;   // this problem's CTAs are [cta0, cta0 + tiles_m * tiles_n * ksplit)
};
struct Params { Problem p[2]; int count; };

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
// raw tile already in the SWIZZLE_128B K-major layout: the split is elementwise, offsets carry over
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

  // which problem / tile / K range
  const int second = (P.count > 1 && (int)blockIdx.x >= P.p[1].cta0) ? 1 : 0;
  const Problem& Q = P.p[second];
  const CUtensorMap* map_a = second ? &map_a1 : &map_a0;
  const CUtensorMap* map_b = second ? &map_b1 : &map_b0;
  const int local = (int)blockIdx.x - Q.cta0;
  const int ks = local % Q.ksplit, tile = local / Q.ksplit;
  const int tn = tile % Q.tiles_n, tm = tile / Q.tiles_n;
  const int nkb_all = (Q.k + KB - 1) / KB;
  const int kb0 = nkb_all * ks / Q.ksplit, kb1 = nkb_all * (ks + 1) / Q.ksplit;
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    prefetch_tmap(map_a);
    prefetch_tmap(map_b);
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

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(smem_u32(&tail->raw_empty[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&tail->raw_full[stage]);
        mbar_arrive_expect_tx(fb, 2 * TILE_BYTES);
        const uint32_t dst_a = base + stage * 2 * TILE_BYTES, dst_b = dst_a + TILE_BYTES;
        if (Q.a_mn) tma_load_2d(dst_a, map_a, fb, tm * TM, kb * KB); else tma_load_2d(dst_a, map_a, fb, kb * KB, tm * TM);
        if (Q.b_mn) tma_load_2d(dst_b, map_b, fb, tn * TN, kb * KB); else tma_load_2d(dst_b, map_b, fb, kb * KB, tn * TN);
        if (++stage == NRAW) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TM, TN);
      int cs = 0; uint32_t cphase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(smem_u32(&tail->conv_full[cs]), cphase);
        tc_fence_after();
        const uint32_t t0 = base + NRAW * 2 * TILE_BYTES + cs * 4 * TILE_BYTES;
        const uint64_t a_hi = make_sw128_kmajor_desc(t0), a_lo = make_sw128_kmajor_desc(t0 + TILE_BYTES);
        const uint64_t b_hi = make_sw128_kmajor_desc(t0 + 2 * TILE_BYTES), b_lo = make_sw128_kmajor_desc(t0 + 3 * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32_ss(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);   // small terms first
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32_ss(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32_ss(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
        umma_commit(smem_u32(&tail->conv_empty[cs]));              // the four split tiles are free when these retire
        if (i == nkb - 1) umma_commit(smem_u32(&tail->acc_full));
        if (++cs == NCONV) { cs = 0; cphase ^= 1; }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // =========================== converters, then epilogue ===========================
    const int t = threadIdx.x - CONV_WARP0 * 32;
    int rs = 0; uint32_t rphase = 0;
    int cs = 0; uint32_t cphase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(smem_u32(&tail->raw_full[rs]), rphase);
      mbar_wait(smem_u32(&tail->conv_empty[cs]), cphase ^ 1);
      const uint8_t* ra = raw_tiles + (size_t)rs * 2 * TILE_BYTES;
      const uint8_t* rb = ra + TILE_BYTES;
      uint8_t* c0 = conv_tiles + (size_t)cs * 4 * TILE_BYTES;
      if (Q.a_mn) convert_mnmajor(ra, c0, c0 + TILE_BYTES, t); else convert_kmajor(ra, c0, c0 + TILE_BYTES, t);
      if (Q.b_mn) convert_mnmajor(rb, c0 + 2 * TILE_BYTES, c0 + 3 * TILE_BYTES, t);
      else convert_kmajor(rb, c0 + 2 * TILE_BYTES, c0 + 3 * TILE_BYTES, t);
      fence_proxy_async();                                         // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&tail->conv_full[cs]));
        mbar_arrive(smem_u32(&tail->raw_empty[rs]));
      }
      if (++rs == NRAW) { rs = 0; rphase ^= 1; }
      if (++cs == NCONV) { cs = 0; cphase ^= 1; }
    }
    // ---- epilogue: thread = TMEM lane = one output row; two warps per lane quarter split the 128 columns
    const int q = warp & 3, h = (warp - CONV_WARP0) >> 2;
    const int row = tm * TM + q * 32 + lane;
    mbar_wait(smem_u32(&tail->acc_full), 0);
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t v[32];
      const int col0 = tn * TN + h * 64 + cc * 32;
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64 + cc * 32), v);
      tmem_ld_wait();
      if (row >= Q.m || col0 >= Q.n) continue;
      const int nvalid = Q.n - col0 >= 32 ? 32 : Q.n - col0;
      if (Q.partial != nullptr) {
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

// operand X(row, k): K-major = x[row * ld + k] -> dims {k, rows}, box {32, 128}, SWIZZLE_128B;
//                    MN-major = x[k * ld + row] -> dims {rows, k}, box {128, 32}, no swizzle.  OOB reads are zero.
static int make_map(CUtensorMap* map, const float* x, int64_t ld, int rows, int k, int mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return FRX_E_DEVICE; }
  cuuint64_t dims[2], strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2], estr[2] = {1, 1};
  if (mn_major) { dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)k; box[0] = TM; box[1] = KB; }
  else { dims[0] = (cuuint64_t)k; dims[1] = (cuuint64_t)rows; box[0] = KB; box[1] = TM; }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
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

// Launch up to two problems in ONE grid.  ksplit > 1 (or keep_partials): raw accumulators go to `partial`
// ([ksplit][m][n] per problem, problem 1 behind problem 0); unless keep_partials, a reduction kernel applies the epilogue.
int gemm3x_launch(cudaStream_t st, const Gemm3xDesc* d, int count, int ksplit, bool keep_partials, float* partial,
                  size_t partial_floats) {
  using namespace g3;
  if (count < 1 || count > 2) { set_error("gemm3x: %d problems (1 or 2 supported)", count); return FRX_E_ARG; }
  for (int i = 0; i < count; ++i)
    if (!gemm3x_supported(d[i])) { set_error("gemm3x: unsupported operand layout (16-byte aligned, ld %% 4 == 0 required)"); return FRX_E_ARG; }
  if (ksplit < 1) ksplit = 1;
  const bool to_partial = ksplit > 1 || keep_partials;
  if (to_partial && (partial == nullptr || partial_floats < gemm3x_partial_floats(d, count, ksplit))) {
    set_error("gemm3x: partial workspace too small");
    return FRX_E_WORKSPACE;
  }
  CUtensorMap maps[4];
  Params P{};
  P.count = count;
  int cta = 0;
  float* pp = partial;
  for (int i = 0; i < count; ++i) {
    int rc = make_map(&maps[2 * i], d[i].a, d[i].lda, d[i].m, d[i].k, d[i].a_mn);
    if (rc) return rc;
    rc = make_map(&maps[2 * i + 1], d[i].b, d[i].ldb, d[i].n, d[i].k, d[i].b_mn);
    if (rc) return rc;
    Problem& Q = P.p[i];
    Q.c = d[i].c; Q.ldc = d[i].ldc;
    Q.partial = to_partial ? pp : nullptr;
    Q.col_scale = d[i].col_scale; Q.col_shift = d[i].col_shift;
    Q.alpha = d[i].alpha; Q.relu = d[i].relu;
    Q.m = d[i].m; Q.n = d[i].n; Q.k = d[i].k; Q.a_mn = d[i].a_mn; Q.b_mn = d[i].b_mn;
    Q.tiles_m = (d[i].m + TM - 1) / TM; Q.tiles_n = (d[i].n + TN - 1) / TN;
    const int nkb = (d[i].k + KB - 1) / KB;
    Q.ksplit = ksplit < nkb ? ksplit : nkb;
    if (to_partial && Q.ksplit != ksplit) { set_error("gemm3x: k = %d too short for a %d-way K split", d[i].k, ksplit); return FRX_E_ARG; }
    Q.cta0 = cta;
    cta += Q.tiles_m * Q.tiles_n * Q.ksplit;
    pp += (size_t)ksplit * (size_t)d[i].m * (size_t)d[i].n;
  }
  if (count == 1) { maps[2] = maps[0]; maps[3] = maps[1]; P.p[1] = P.p[0]; P.p[1].cta0 = cta; }
  static bool attr_set = false;
  if (!attr_set) {
    FRX_CUDA(cudaFuncSetAttribute(gemm3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  gemm3x_kernel<<<cta, NUM_THREADS, SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], P);
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
  return ks > 1 ? frx::gemm3x_partial_floats(&d, 1, ks) * sizeof(float) + 256 : 256;
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
  if (ks > 1 && (workspace == nullptr || workspace_bytes < frx_linear_workspace_bytes(m, n, k))) {
    set_error("frx_linear: workspace %zu bytes, need %zu", workspace_bytes, frx_linear_workspace_bytes(m, n, k));
    return FRX_E_WORKSPACE;
  }
  return gemm3x_launch((cudaStream_t)stream, &d, 1, ks, false, reinterpret_cast<float*>(workspace),
                       workspace_bytes / sizeof(float));
}

}  // extern "C"
