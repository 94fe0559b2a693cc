// Exact mode / distinct count of numeric columns: batched LSD radix sort of the
// order-preserving keys of the non-null values + a run-length summary of the sorted keys.
//
// Replaces, for numeric columns, the per-column `groupBy(col).count().orderBy(desc).limit(1)`
// Spark jobs of mode_computation and the countDistinct aggregation of
// uniqueCount_computation (reference /root/reference/src/main/anovos/data_analyzer/
// stats_generator.py:386-401 and :611).  String columns never come here (their dictionary
// code histograms already hold the group counts).
//
// Pipeline (all columns of the call advance together, grid.y = column):
//   pack     values -> keys, nulls dropped, exact zeros counted instead of sorted (block compaction: the order
//            before a sort is irrelevant), key count per column kept on the device;
//   8-bit LSD passes: tile histogram -> per-column exclusive scan -> stable scatter (8-ballot peer ranking,
//            shared-memory reorder); a pass in which one digit holds every key is skipped (device-side
//            decision taken from the scanned table, no host sync);
//   runs     per-thread run summaries of 16 consecutive sorted keys (head count, open prefix / suffix run,
//            longest closed run), combined with an associative operator per warp, per tile and per column;
//            the merge kernel splices the zero run back and reads the requested order statistics.
// Counting is integer everywhere => deterministic.  Ties for the mode resolve to the
// smallest value (the reference's choice is arbitrary, stats_generator.py:358).
#include <stdlib.h>

#include "common.cuh"
#include "hll_hash.cuh"

namespace anv {

#ifndef ANV_SCAT_MINB
#define ANV_SCAT_MINB 2          // resident scatter CTAs per SM the register budget is tuned for
#endif
constexpr int SORT_TILE = 4096;  // keys per CTA
constexpr int SCAT_THREADS = 512;  // the scatter kernel runs 16 warps x 8 rounds of 32 keys
constexpr int SCAT_WARPS = SCAT_THREADS / 32;

template <typename K, typename T> __device__ __forceinline__ K make_key(T x);
template <> __device__ __forceinline__ uint32_t make_key<uint32_t, float>(float x) {
  x += 0.0f;
  const uint32_t u = __float_as_uint(x);
  return (x != x) ? 0xFFFFFFFFu : ((u & 0x80000000u) ? ~u : (u | 0x80000000u));
}
template <> __device__ __forceinline__ uint32_t make_key<uint32_t, int32_t>(int32_t x) { return (uint32_t)x ^ 0x80000000u; }
template <> __device__ __forceinline__ uint64_t make_key<uint64_t, double>(double x) {
  x += 0.0;
  const uint64_t u = (uint64_t)__double_as_longlong(x);
  return (x != x) ? ~0ull : ((u >> 63) ? ~u : (u | (1ull << 63)));
}
template <> __device__ __forceinline__ uint64_t make_key<uint64_t, int64_t>(int64_t x) { return (uint64_t)x ^ (1ull << 63); }
template <> __device__ __forceinline__ uint64_t make_key<uint64_t, float>(float x) { return (uint64_t)make_key<uint32_t, float>(x) << 32; }
template <> __device__ __forceinline__ uint64_t make_key<uint64_t, int32_t>(int32_t x) { return (uint64_t)make_key<uint32_t, int32_t>(x) << 32; }
template <> __device__ __forceinline__ uint32_t make_key<uint32_t, double>(double) { return 0; }   // never used
template <> __device__ __forceinline__ uint32_t make_key<uint32_t, int64_t>(int64_t) { return 0; }

__device__ __forceinline__ double sorted_key_to_double(uint64_t k, int dtype) {
  switch (dtype) {
    case ANV_F32: {
      uint32_t u = (uint32_t)(k >> 32);
      u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
      return (double)__uint_as_float(u);
    }
    case ANV_I32: return (double)(int32_t)((uint32_t)(k >> 32) ^ 0x80000000u);
    case ANV_F64: {
      const uint64_t u = (k >> 63) ? (k & ~(1ull << 63)) : ~k;
      return __longlong_as_double((long long)u);
    }
    default: return (double)(int64_t)(k ^ (1ull << 63));
  }
}

// Spark's hash of the VALUE a sorted key stands for (the HLL++ by-product of the run summaries): undo the order-preserving
// transform; every NaN has the key ~0 and hashes as the canonical NaN, like Spark's floatToIntBits / doubleToLongBits.
template <typename K> __device__ __forceinline__ uint64_t spark_hash_of_key(K k, int dtype);
template <> __device__ __forceinline__ uint64_t spark_hash_of_key<uint32_t>(uint32_t k, int dtype) {
  if (dtype == ANV_I32) return xxh64_int(k ^ 0x80000000u);
  const uint32_t u = (k == 0xFFFFFFFFu) ? 0x7fc00000u : ((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
  return xxh64_int(u);
}
template <> __device__ __forceinline__ uint64_t spark_hash_of_key<uint64_t>(uint64_t k, int dtype) {
  switch (dtype) {
    case ANV_F32: return spark_hash_of_key<uint32_t>((uint32_t)(k >> 32), ANV_F32);
    case ANV_I32: return spark_hash_of_key<uint32_t>((uint32_t)(k >> 32), ANV_I32);
    case ANV_I64: return xxh64_long(k ^ (1ull << 63));
    default: {
      const uint64_t u = (k == ~0ull) ? 0x7ff8000000000000ull : ((k >> 63) ? (k & ~(1ull << 63)) : ~k);
      return xxh64_long(u);
    }
  }
}

struct ColState {               // one per column, in the workspace
  unsigned long long n_valid;   // filled by pack_kernel: non-null, NONZERO values = keys that are sorted
  unsigned long long n_zero;    // filled by pack_kernel: non-null values equal to 0 (kept out of the sort, see pack)
  int cur;                      // which ping-pong buffer holds the current order
  int src[8];                   // per pass: source buffer
  int skip[8];                  // per pass: digit constant -> no scatter
  int error;                    // look-back spin limit hit (never expected): the host raises instead of hanging
};

template <typename K> struct TileSummary {
  K first_key, last_key, best_key;
  uint32_t n, prefix_len, suffix_len, best_len, heads_inside;
};

template <typename K> struct SortParams {
  const anv_column_t* cols;
  int n_cols;
  int64_t n_rows;
  int64_t stride;               // keys per column in each buffer
  int n_tiles;                  // ceil(n_rows / SORT_TILE)
  K* buf[2];
  ColState* state;
  uint32_t* tile_hist;          // [n_cols][256][n_tiles]  (digit-major)
  TileSummary<K>* summ;         // [n_cols][n_tiles]
  int pass;
  // one-sweep passes (decoupled look-back): no tile-histogram and no scan kernel
  uint32_t* ghist;              // [n_cols][sizeof(K)][256]  digit counts of the whole column, all passes, taken by pack_kernel
  uint32_t* gbase;              // [n_cols][sizeof(K)][256]  their exclusive scans
  unsigned long long* status;   // [n_cols][n_tiles][256]   (epoch << 56 | kind << 54 | count): tile aggregates / inclusive prefixes
  uint32_t* ticket;             // [n_cols][sizeof(K)]       tile ids are handed out in arrival order
  // optional by-product: HyperLogLog++ registers (Spark's approx_count_distinct) from the DISTINCT sorted keys
  int hll_p;                    // 0 = off; 4..12
  uint32_t* hll_regs;           // [n_cols][1 << hll_p], zeroed by the host wrapper
};
constexpr int PACK_TPC = 8;     // tiles per pack CTA (amortises the flush of the digit histograms)

// byte `pass` of the key: ONE PRMT (selector nibble 0 = the byte, the other result bytes come from the zero operand)
// instead of a shift and a mask - the digit is extracted three times per key and pass in the scatter.
template <typename K> __device__ __forceinline__ uint32_t digit_of(K k, int pass);
template <> __device__ __forceinline__ uint32_t digit_of<uint32_t>(uint32_t k, int pass) {
  return __byte_perm(k, 0u, 0x4440u | (uint32_t)pass);
}
template <> __device__ __forceinline__ uint32_t digit_of<uint64_t>(uint64_t k, int pass) {
  const uint32_t half = (pass & 4) ? (uint32_t)(k >> 32) : (uint32_t)k;
  return __byte_perm(half, 0u, 0x4440u | (uint32_t)(pass & 3));
}

// ---- pack: values -> keys, nulls dropped --------------------------------------------------
// One CTA per 4096-row tile: 128-bit loads, block-level compaction (one atomicAdd per CTA
// reserves the output range; the order before a sort is irrelevant), 64-byte runs per thread.
// Exact zeros are COUNTED here instead of being sorted: sparse / zero-inflated columns (70 % zeros in the
// benchmark's fourth family, > 90 % in the income dataset's capital-gain / capital-loss) then sort only their
// nonzero values; run_merge_kernel splices the zero run back into mode, distinct count and ranks.
template <typename K, typename T>
__device__ __forceinline__ void pack_tile(const SortParams<K>& P, const anv_column_t& col, int c, const int64_t tile, uint32_t* s_warp,
                                          unsigned long long* s_base, K* sk, uint32_t (*s_dh)[256]) {
  constexpr int VEC = Traits<T>::VEC;
  constexpr int PER = SORT_TILE / ANV_BLOCK;  // 16 rows per thread, contiguous
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = tile * SORT_TILE;
  const int n_tile = (int)min((int64_t)SORT_TILE, P.n_rows - r0);
  const T* __restrict__ data = reinterpret_cast<const T*>(col.data) + r0;
  const uint32_t* __restrict__ vbits = col.validity;
  const int row0 = tid * PER;
  K keys[PER];
  uint32_t okmask = 0;
  if (row0 + PER <= n_tile) {
    uint32_t vb = 0xFFFFu;
    if (vbits) {
      const int64_t g = r0 + row0;  // multiple of 16
      vb = (__ldg(vbits + (g >> 5)) >> (g & 31)) & 0xFFFFu;
    }
    okmask = vb;
    const uint4* p = reinterpret_cast<const uint4*>(data + row0);
#pragma unroll
    for (int v = 0; v < PER / VEC; ++v) {
      const uint4 q = ldg_stream(p + v);
      T e[VEC];
      unpack<T>(q, e);
#pragma unroll
      for (int i = 0; i < VEC; ++i) keys[v * VEC + i] = make_key<K, T>(e[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int row = row0 + i;
      bool ok = row < n_tile;
      keys[i] = 0;
      if (ok) {
        if (vbits) { const int64_t g = r0 + row; ok = (vbits[g >> 5] >> (g & 31)) & 1u; }
        keys[i] = make_key<K, T>(data[row]);
      }
      okmask |= ok ? (1u << i) : 0u;
    }
  }
  {  // take the zeros out (0.0 and -0.0 share one key; integers: 0)
    constexpr K ZERO_KEY = (K)1 << (sizeof(K) * 8 - 1);
    uint32_t zmask = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) zmask |= (keys[i] == ZERO_KEY) ? (1u << i) : 0u;
    zmask &= okmask;
    okmask &= ~zmask;
    const uint32_t wz = __reduce_add_sync(ANV_FULL, (uint32_t)__popc(zmask));
    if (lane == 0 && wz) atomicAdd(&P.state[c].n_zero, (unsigned long long)wz);
  }
  // block exclusive scan of the per-thread valid counts
  const uint32_t mine = __popc(okmask);
  uint32_t inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(ANV_FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (tid == 0) {
    uint32_t acc = 0;
    for (int w = 0; w < ANV_WARPS; ++w) { const uint32_t t = s_warp[w]; s_warp[w] = acc; acc += t; }
    *s_base = acc ? atomicAdd(&P.state[c].n_valid, (unsigned long long)acc) : 0ull;
  }
  __syncthreads();
  // stage the compacted keys in shared memory, then write them out coalesced
  uint32_t o = s_warp[warp] + (inc - mine);
#pragma unroll
  for (int i = 0; i < PER; ++i)
    if ((okmask >> i) & 1u) sk[o++] = keys[i];
  __shared__ uint32_t s_total;
  if (tid == ANV_BLOCK - 1) s_total = s_warp[warp] + inc;
  __syncthreads();
  K* __restrict__ out = P.buf[0] + (size_t)c * P.stride + *s_base;
  const uint32_t total = s_total;
  // copy-out + the digit histograms of EVERY pass from the staged keys (the one-sweep passes need the column-wide digit counts
  // before the first scatter; taking them here replaces one full read of the keys per pass)
  if (P.ghist) {
    for (uint32_t i = tid; i < total; i += ANV_BLOCK) {
      const K k = sk[i];
      out[i] = k;
#pragma unroll
      for (int ps = 0; ps < (int)sizeof(K); ++ps) atomicAdd(&s_dh[ps][digit_of(k, ps)], 1u);
    }
  } else {
    for (uint32_t i = tid; i < total; i += ANV_BLOCK) out[i] = sk[i];
  }
  __syncthreads();   // sk / s_warp are reused by the CTA's next tile
}

// Strided variant (ANV_PACK_STRIDED): thread t takes rows t, t + 256, ... of the tile (coalesced scalar loads, the bitmap
// word of a warp's 32 rows is one broadcast load), a ballot per round ranks the surviving keys of the warp, and the staging
// stores land on CONSECUTIVE shared-memory words per warp.  The contiguous variant above (16 consecutive rows per thread)
// writes its staging stores 16 words apart - two banks per warp, 16-way conflicts (ncu: mio_throttle is its top stall).
template <typename T> __device__ __forceinline__ T ld_stream_scalar(const T* p);
template <> __device__ __forceinline__ float ld_stream_scalar<float>(const float* p) {
  float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
template <> __device__ __forceinline__ int32_t ld_stream_scalar<int32_t>(const int32_t* p) {
  int32_t v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
template <> __device__ __forceinline__ double ld_stream_scalar<double>(const double* p) {
  double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}
template <> __device__ __forceinline__ int64_t ld_stream_scalar<int64_t>(const int64_t* p) {
  long long v; asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p)); return (int64_t)v;
}

template <typename K, typename T>
__device__ __forceinline__ void pack_tile_strided(const SortParams<K>& P, const anv_column_t& col, int c, const int64_t tile, uint32_t* s_warp,
                                                  unsigned long long* s_base, K* sk, uint32_t (*s_dh)[256]) {
  constexpr int PER = SORT_TILE / ANV_BLOCK;        // 16 rounds of 256 rows
  constexpr int WREG = SORT_TILE / ANV_WARPS;       // a warp stages at most 16 x 32 keys: its own 512 slots of sk
  constexpr K ZERO_KEY = (K)1 << (sizeof(K) * 8 - 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = tile * SORT_TILE;              // multiple of 4096: the tile starts on a bitmap word
  const int n_tile = (int)min((int64_t)SORT_TILE, P.n_rows - r0);
  const T* __restrict__ data = reinterpret_cast<const T*>(col.data) + r0 + tid;
  const uint32_t* __restrict__ vw = col.validity ? col.validity + (r0 >> 5) + warp : nullptr;   // round i: word i * 8 + warp, bit = lane
  const uint32_t lbit = 1u << lane, lt = lbit - 1u;
  K* const mine = sk + warp * WREG;
  const bool full = n_tile == SORT_TILE;
  T x[PER];
  // the warp's 16 bitmap words (one per round), lane i holds round i's: one load per lane instead of 16 broadcast loads
  uint32_t wv = ANV_FULL;
  if (vw && lane < PER && (full || lane * ANV_BLOCK + warp * 32 < n_tile)) wv = __ldg(vw + lane * (ANV_BLOCK / 32));
#pragma unroll
  for (int i = 0; i < PER; ++i) {                   // all loads of the tile in flight at once
    const bool in = full || (i * ANV_BLOCK + tid) < n_tile;
    x[i] = in ? ld_stream_scalar<T>(data + i * ANV_BLOCK) : (T)0;
  }
  uint32_t wcount = 0, nzero = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const K k = make_key<K, T>(x[i]);
    bool ok = (__shfl_sync(ANV_FULL, wv, i) & lbit) != 0u;
    if (!full) ok = ok && (i * ANV_BLOCK + tid) < n_tile;
    const bool zero = ok && k == ZERO_KEY;          // 0.0 and -0.0 share one key; integers: 0
    nzero += zero ? 1u : 0u;
    ok = ok && !zero;
    const uint32_t bal = __ballot_sync(ANV_FULL, ok);
    if (ok) mine[wcount + __popc(bal & lt)] = k;    // consecutive lanes, consecutive words
    wcount += __popc(bal);                          // warp-uniform
  }
  const uint32_t wz = __reduce_add_sync(ANV_FULL, nzero);
  if (lane == 0 && wz) atomicAdd(&P.state[c].n_zero, (unsigned long long)wz);
  if (lane == 0) s_warp[warp] = wcount;
  __syncthreads();
  if (tid == 0) {
    uint32_t acc = 0;
    for (int ww = 0; ww < ANV_WARPS; ++ww) { const uint32_t t = s_warp[ww]; s_warp[ww] = acc; acc += t; }
    *s_base = acc ? atomicAdd(&P.state[c].n_valid, (unsigned long long)acc) : 0ull;
  }
  __syncthreads();
  // every warp copies its own staged keys out (the order before a sort is irrelevant): contiguous, coalesced
  K* __restrict__ out = P.buf[0] + (size_t)c * P.stride + *s_base + s_warp[warp];
  if (P.ghist) {
    for (uint32_t j = lane; j < wcount; j += 32) {
      const K k = mine[j];
      out[j] = k;
#pragma unroll
      for (int ps = 0; ps < (int)sizeof(K); ++ps) atomicAdd(&s_dh[ps][digit_of(k, ps)], 1u);
    }
  } else {
    for (uint32_t j = lane; j < wcount; j += 32) out[j] = mine[j];
  }
  __syncthreads();   // sk / s_warp / s_base are reused by the CTA's next tile
}

#ifndef ANV_PACK_STRIDED
#define ANV_PACK_STRIDED 1   // measured (100 M x 12 float32, whole sort call): contiguous 27.29 ms, strided 26.38 ms, identical results
#endif
#if ANV_PACK_STRIDED
#define ANV_PACK_TILE pack_tile_strided
#else
#define ANV_PACK_TILE pack_tile
#endif

#ifndef ANV_PACK_MINB
#define ANV_PACK_MINB 6        // 40 registers, no spills for 32-bit keys (1: 26.64 ms, 5: 26.44, 6: 26.38); 64-bit keys: 5
#endif
template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK, (sizeof(K) == 8 && ANV_PACK_MINB > 5) ? 5 : ANV_PACK_MINB) pack_kernel(const SortParams<K> P) {
  __shared__ uint32_t s_warp[ANV_WARPS];
  __shared__ unsigned long long s_base;
  __shared__ K sk[SORT_TILE];
  __shared__ uint32_t s_dh[sizeof(K)][256];
  const int c = blockIdx.y, tid = threadIdx.x;
  const anv_column_t col = P.cols[c];
#pragma unroll
  for (int ps = 0; ps < (int)sizeof(K); ++ps) s_dh[ps][tid] = 0;
  __syncthreads();
  for (int t = 0; t < PACK_TPC; ++t) {
    const int64_t tile = (int64_t)blockIdx.x * PACK_TPC + t;
    if (tile * SORT_TILE >= P.n_rows) break;
    switch (col.dtype) {
      case ANV_F32: ANV_PACK_TILE<K, float>(P, col, c, tile, s_warp, &s_base, sk, s_dh); break;
      case ANV_I32: ANV_PACK_TILE<K, int32_t>(P, col, c, tile, s_warp, &s_base, sk, s_dh); break;
      case ANV_F64: if (sizeof(K) == 8) ANV_PACK_TILE<K, double>(P, col, c, tile, s_warp, &s_base, sk, s_dh); break;
      case ANV_I64: if (sizeof(K) == 8) ANV_PACK_TILE<K, int64_t>(P, col, c, tile, s_warp, &s_base, sk, s_dh); break;
      default: break;
    }
  }
  if (!P.ghist) return;
  __syncthreads();
  uint32_t* g = P.ghist + (size_t)c * sizeof(K) * 256;
#pragma unroll
  for (int ps = 0; ps < (int)sizeof(K); ++ps) {
    const uint32_t v = s_dh[ps][tid];
    if (v) atomicAdd(&g[ps * 256 + tid], v);
  }
}

// Column-wide exclusive digit offsets of every pass + which passes are no-ops (one digit holds every key) + the ping-pong
// buffer each pass reads: everything the one-sweep passes need is known after the pack kernel.
template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK) sort_bases_kernel(const SortParams<K> P) {
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ColState& S = P.state[c];
  const unsigned long long n = S.n_valid;
  __shared__ uint32_t wsum[ANV_WARPS];
  __shared__ int s_skip[sizeof(K)];
  if (tid < (int)sizeof(K)) s_skip[tid] = (n == 0) ? 1 : 0;
  __syncthreads();
  for (int ps = 0; ps < (int)sizeof(K); ++ps) {
    const uint32_t v = P.ghist[((size_t)c * sizeof(K) + ps) * 256 + tid];
    if (n > 0 && (unsigned long long)v == n) s_skip[ps] = 1;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(ANV_FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < ANV_WARPS; ++w) woff += (w < warp) ? wsum[w] : 0u;
    P.gbase[((size_t)c * sizeof(K) + ps) * 256 + tid] = woff + inc - v;
    __syncthreads();
  }
  if (tid == 0) {
    int cur = 0;
    for (int ps = 0; ps < (int)sizeof(K); ++ps) {
      S.src[ps] = cur;
      S.skip[ps] = s_skip[ps];
      if (!s_skip[ps]) cur ^= 1;
    }
    S.cur = cur;
  }
}

// Lanes of `act` holding the same 8-bit digit.  Per bit: test, ballot of "my bit is set", and ONE three-input logic op
// m &= ballot ^ (my bit ? 0 : ~0)  - keep the lanes whose bit equals mine (4 SASS instructions per bit + the select).
#define ANV_PEER_BIT(B)                                                                                       \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t, bal, s;\n\tand.b32 t, %1, " #B ";\n\tsetp.ne.u32 p, t, 0;\n\t"     \
               "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\tselp.b32 s, 0, -1, p;\n\t"                               \
               "lop3.b32 %0, %0, bal, s, 0x60;\n\t}"                                                          \
               : "+r"(m) : "r"(d))
__device__ __forceinline__ uint32_t peers8(uint32_t d, uint32_t act) {
  uint32_t m = act;
  ANV_PEER_BIT(1); ANV_PEER_BIT(2); ANV_PEER_BIT(4); ANV_PEER_BIT(8);
  ANV_PEER_BIT(16); ANV_PEER_BIT(32); ANV_PEER_BIT(64); ANV_PEER_BIT(128);
  return m;
}
#undef ANV_PEER_BIT


// ---- pass step 1: per-tile digit histogram ------------------------------------------------------
// Plain shared-memory atomics (hardware handles same-address lanes far faster than a
// match_any pre-aggregation: measured 4-5x on B200).
#ifndef ANV_HIST_TPC
#define ANV_HIST_TPC 4
#endif
// tiles per tile-histogram CTA.  Measured (c2 / c3 sort call, ms): 1 tile 11.36 / 317.3, 4 tiles 11.19 / 314.8 - the two dependent
// loads that start a CTA (column state, then keys) are paid once per 64 KB.  The same knob on the scatter (2 tiles: 12.42 /
// 354.6) and the run summaries (4 tiles: 11.29 / 316.8) does not pay and stays at 1.
constexpr int HIST_TPC = ANV_HIST_TPC;
template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK) sort_hist_kernel(const SortParams<K> P) {
  const int c = blockIdx.y, tid = threadIdx.x;
  const ColState& S = P.state[c];
  const int64_t n = (int64_t)S.n_valid;
  const K* __restrict__ col_keys = (S.cur ? P.buf[1] : P.buf[0]) + (size_t)c * P.stride;
  __shared__ uint32_t h[256];
  const int sh = P.pass * 8;
  for (int tt = 0; tt < HIST_TPC; ++tt) {
    const int tile = blockIdx.x * HIST_TPC + tt;
    if (tile >= P.n_tiles) break;
    const int64_t t0 = (int64_t)tile * SORT_TILE;
    h[tid] = 0;
    __syncthreads();
    if (t0 < n) {
      const K* __restrict__ keys = col_keys + t0;
      const int nt = (int)min((int64_t)SORT_TILE, n - t0);
      constexpr int KV = 16 / sizeof(K);  // keys per 128-bit load
      const int nvec = nt / KV;
      const uint4* __restrict__ kv = reinterpret_cast<const uint4*>(keys);
      for (int j = tid; j < nvec; j += ANV_BLOCK) {
        const uint4 q = kv[j];
        if (sizeof(K) == 4) {
          atomicAdd(&h[(q.x >> sh) & 0xFFu], 1u); atomicAdd(&h[(q.y >> sh) & 0xFFu], 1u);
          atomicAdd(&h[(q.z >> sh) & 0xFFu], 1u); atomicAdd(&h[(q.w >> sh) & 0xFFu], 1u);
        } else {
          const uint64_t k0 = ((uint64_t)q.y << 32) | q.x, k1 = ((uint64_t)q.w << 32) | q.z;
          atomicAdd(&h[(uint32_t)(k0 >> sh) & 0xFFu], 1u); atomicAdd(&h[(uint32_t)(k1 >> sh) & 0xFFu], 1u);
        }
      }
      for (int i = nvec * KV + tid; i < nt; i += ANV_BLOCK) atomicAdd(&h[digit_of(keys[i], P.pass)], 1u);
    }
    __syncthreads();
    P.tile_hist[((size_t)c * 256 + tid) * P.n_tiles + tile] = h[tid];
    __syncthreads();
  }
}

// ---- pass step 2: exclusive scan of [256][n_tiles] per column + skip decision ------------------------------------
// Two launches with one CTA per (digit, column) - 256 x n_cols CTAs instead of n_cols (a batch of 15-50 columns left most of
// the 148 SMs idle while a single CTA per column walked 6 M entries at c3):
//   sort_totals_kernel   digit_total[c][d] = sum over tiles of tile_hist[c][d][*]
//   sort_scan_kernel     base(d) = sum of the totals of the smaller digits (256 values, one warp scan per CTA), then the exclusive
//                        scan of the digit's own tile counts on top of it; the CTA of digit 0 also takes the device-side
//                        "one digit holds every key -> skip the pass" decision.
template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK) sort_totals_kernel(const SortParams<K> P, uint32_t* __restrict__ totals) {
  const int d = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
  const uint32_t* __restrict__ a = P.tile_hist + ((size_t)c * 256 + d) * P.n_tiles;
  uint32_t acc = 0;
  const int n4 = P.n_tiles & ~3;
  if ((((size_t)c * 256 + d) * P.n_tiles & 3) == 0) {
    for (int i = tid * 4; i < n4; i += ANV_BLOCK * 4) { const uint4 q = *reinterpret_cast<const uint4*>(a + i); acc += q.x + q.y + q.z + q.w; }
    for (int i = n4 + tid; i < P.n_tiles; i += ANV_BLOCK) acc += a[i];
  } else {
    for (int i = tid; i < P.n_tiles; i += ANV_BLOCK) acc += a[i];
  }
  acc = __reduce_add_sync(ANV_FULL, acc);
  __shared__ uint32_t w[ANV_WARPS];
  if ((tid & 31) == 0) w[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
    for (int i = 0; i < ANV_WARPS; ++i) t += w[i];
    totals[(size_t)c * 256 + d] = t;
  }
}

template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK) sort_scan_kernel(const SortParams<K> P, const uint32_t* __restrict__ totals) {
  const int d = blockIdx.x, c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ColState& S = P.state[c];
  const unsigned long long n = S.n_valid;
  __shared__ uint32_t s_base, s_skip;
  __shared__ uint32_t wsum[ANV_WARPS + 1];
  if (tid == 0) s_skip = 0;
  __syncthreads();
  {  // base of this digit + skip decision from the 256 totals of the column (each thread owns one digit)
    const uint32_t t = totals[(size_t)c * 256 + tid];
    if (n > 0 && (unsigned long long)t == n) s_skip = 1;          // (benign race: every writer writes 1)
    uint32_t below = (tid < d) ? t : 0u;
    below = __reduce_add_sync(ANV_FULL, below);
    if (lane == 0) wsum[warp] = below;
    __syncthreads();
    if (tid == 0) {
      uint32_t b = 0;
      for (int i = 0; i < ANV_WARPS; ++i) b += wsum[i];
      s_base = b;
    }
    __syncthreads();
  }
  if (d == 0 && tid == 0) {
    const int skip = (n == 0) ? 1 : (int)s_skip;
    S.src[P.pass] = S.cur;
    S.skip[P.pass] = skip;
    if (!skip) S.cur ^= 1;
  }
  if (n == 0) return;
  uint32_t* a = P.tile_hist + ((size_t)c * 256 + d) * P.n_tiles;
  uint32_t carry = s_base;
  constexpr int PER = 8;
  for (int base = 0; base < P.n_tiles; base += ANV_BLOCK * PER) {
    uint32_t v[PER], run = 0;
    const int i0 = base + tid * PER;
#pragma unroll
    for (int k = 0; k < PER; ++k) { v[k] = (i0 + k < P.n_tiles) ? a[i0 + k] : 0u; run += v[k]; }
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(ANV_FULL, inc, o);
      if (lane >= o) inc += t;
    }
    __syncthreads();                       // wsum of the previous chunk has been consumed
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < ANV_WARPS; ++w) { const uint32_t t = wsum[w]; woff += (w < warp) ? t : 0u; tot += t; }
    uint32_t ex = carry + woff + inc - run;
#pragma unroll
    for (int k = 0; k < PER; ++k) { if (i0 + k < P.n_tiles) a[i0 + k] = ex; ex += v[k]; }
    carry += tot;
  }
}

// ---- pass step 3: stable scatter ---------------------------------------------------------------
// Ranking: warp w owns 512 consecutive keys (16 rounds of 32); the lanes holding the same digit
// are found with 8 ballots (cheaper than MATCH.ANY on sm_100a), all 16 rounds' loads and peer
// masks are computed up front (independent), then the warp-private digit counters are advanced
// round by round.  The tile is then REORDERED IN SHARED MEMORY into digit order, so the global
// writes are coalesced runs (full 32-byte sectors) instead of 4-byte scatters.
template <typename K> struct ScatShared {   // declared ONCE in the kernel (statics in the templated body would be replicated)
  uint16_t wcnt[SCAT_WARPS][256];             // <= 4096 keys per tile: 16 bits are enough
  uint32_t gbase[256];
  uint32_t wtot[8];
  K sk[SORT_TILE + 1];                        // + the spare slot of the branch-free placement
};

constexpr unsigned long long LB_AGG = 1ull << 54, LB_PREFIX = 2ull << 54, LB_VALUE = (1ull << 54) - 1ull;
constexpr int LB_SPIN_LIMIT = 1 << 20;

// Decoupled look-back of ONE digit: add up the aggregates of the preceding tiles until one that already knows its inclusive
// prefix, then publish this tile's inclusive prefix.  (Not inlined: keeps the spin loop out of the scatter's control flow.)
__device__ __noinline__ unsigned long long lookback_exclusive(volatile unsigned long long* status, int tile, int d, unsigned long long epoch,
                                                             unsigned long long total, int* error) {
  unsigned long long excl = 0;
  for (int t = tile - 1; t >= 0; --t) {
    unsigned long long v = status[(size_t)t * 256 + d];
    int spins = 0;
    while ((v >> 56) != (epoch >> 56)) {
      // never expected (tiles start in ticket order); once one look-back has given up every other one follows at once
      if (++spins > LB_SPIN_LIMIT || ((spins & 255) == 0 && *reinterpret_cast<volatile int*>(error))) { *error = 1; return excl; }
      v = status[(size_t)t * 256 + d];
    }
    excl += v & LB_VALUE;
    if (v & LB_PREFIX) break;
  }
  status[(size_t)tile * 256 + d] = epoch | LB_PREFIX | (excl + total);
  return excl;
}

// store through the global window (the opaque base pointer below would otherwise compile to a generic ST)
__device__ __forceinline__ void st_global_key(uint32_t* p, uint32_t v) { asm volatile("st.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_global_key(uint64_t* p, uint64_t v) { asm volatile("st.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

template <typename K, bool FULL, bool LOOKBACK>
__device__ __forceinline__ void scatter_tile(const SortParams<K>& P, const ColState& S, const int c, const int tile, const int64_t t0,
                                             const int nt_in, ScatShared<K>& SH) {
  auto& wcnt = SH.wcnt;
  auto& gbase = SH.gbase;
  auto& wtot = SH.wtot;
  auto& sk = SH.sk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = FULL ? SORT_TILE : nt_in;   // a full tile needs no bounds logic (all but the last tile of a column)
  const int src = S.src[P.pass];
  const K* __restrict__ in = (src ? P.buf[1] : P.buf[0]) + (size_t)c * P.stride + t0;
  K* __restrict__ out = (src ? P.buf[0] : P.buf[1]) + (size_t)c * P.stride;
  for (int i = tid; i < SCAT_WARPS * 256 / 2; i += SCAT_THREADS) reinterpret_cast<uint32_t*>(&wcnt[0][0])[i] = 0;
  if (!LOOKBACK && tid < 256) gbase[tid] = P.tile_hist[((size_t)c * 256 + tid) * P.n_tiles + tile];
  constexpr int WR = SORT_TILE / SCAT_WARPS / 32;  // 8 rounds per warp
  K key[WR];
  uint32_t peers[WR];
  const int w0 = warp * (SORT_TILE / SCAT_WARPS);
#pragma unroll
  for (int r = 0; r < WR; ++r) {
    const int i = w0 + r * 32 + lane;
    key[r] = (FULL || i < nt) ? in[i] : (K)0;
  }
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < WR; ++r) {
    const bool ok = FULL || (w0 + r * 32 + lane) < nt;
    const uint32_t act = FULL ? ANV_FULL : __ballot_sync(ANV_FULL, ok);
    const uint32_t d = digit_of(key[r], P.pass);
    const uint32_t m = peers8(d, act);
    peers[r] = ok ? m : 0u;
  }
  __syncthreads();
  uint32_t pos[WR];
#pragma unroll
  for (int r = 0; r < WR; ++r) {
    const uint32_t m = peers[r];
    const uint32_t d = digit_of(key[r], P.pass);
    uint32_t before = 0;
    if (m) before = wcnt[warp][d];
    const uint32_t lower = m & lt;               // peers in lower lanes: none <=> this lane leads its group
    pos[r] = before + __popc(lower);
    __syncwarp();
    if (m && lower == 0) wcnt[warp][d] = (uint16_t)(before + __popc(m));
    __syncwarp();
  }
  __syncthreads();
  uint32_t total = 0;
  if (tid < 256) {  // exclusive prefix over warps for digit `tid`, tile total of the digit
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < SCAT_WARPS; ++w) { const uint32_t t = wcnt[w][tid]; wcnt[w][tid] = (uint16_t)acc; acc += t; }
    total = acc;
  }
  const unsigned long long epoch = (unsigned long long)(P.pass + 1) << 56;
  volatile unsigned long long* const status = LOOKBACK ? P.status + ((size_t)c * P.n_tiles) * 256 : nullptr;
  if (LOOKBACK && tid < 256)   // publish this tile's digit counts at once: its successors only need these to move on
    status[(size_t)tile * 256 + tid] = epoch | (tile == 0 ? LB_PREFIX : LB_AGG) | (unsigned long long)total;
  uint32_t ds_keep = 0;
  {  // exclusive scan of the 256 digit totals -> start of each digit inside the reordered tile
    uint32_t inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(ANV_FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31 && warp < 8) wtot[warp] = inc;
    __syncthreads();
    if (tid < 256) {
      uint32_t woff = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) woff += (w < warp) ? wtot[w] : 0u;
      const uint32_t ds = woff + inc - total;
      // fold the digit's start into the per-warp offsets (one lookup when placing) and keep a single
      // "global base minus tile start" table for the copy-out (one lookup per key there as well)
#pragma unroll
      for (int w = 0; w < SCAT_WARPS; ++w) wcnt[w][tid] = (uint16_t)(wcnt[w][tid] + ds);
      if (LOOKBACK) ds_keep = ds;
      else gbase[tid] -= ds;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < WR; ++r) {   // branch-free: lanes past the end of a partial tile write to the spare slot
    const uint32_t d = digit_of(key[r], P.pass);
    const uint32_t at = wcnt[warp][d] + pos[r];
    sk[(FULL || peers[r]) ? at : (uint32_t)SORT_TILE] = key[r];
  }
  if (LOOKBACK && tid < 256) {
    const unsigned long long excl = lookback_exclusive(status, tile, tid, epoch, (unsigned long long)total, &P.state[c].error);
    gbase[tid] = P.gbase[((size_t)c * sizeof(K) + P.pass) * 256 + tid] + (uint32_t)excl - ds_keep;
  }
  __syncthreads();
  // the column's output base as ONE 64-bit register pair (opaque to the compiler: it otherwise re-derives
  // c * stride + index in 64 bits for every key - 4 instructions per key of the copy-out)
  K* outc = out;
  asm volatile("" : "+l"(outc));
  if (FULL) {
#pragma unroll
    for (int j = 0; j < SORT_TILE / SCAT_THREADS; ++j) {
      const int p = tid + j * SCAT_THREADS;
      const K k = sk[p];
      st_global_key(outc + (gbase[digit_of(k, P.pass)] + (uint32_t)p), k);
    }
  } else {
    for (int p = tid; p < nt; p += SCAT_THREADS) {
      const K k = sk[p];
      st_global_key(outc + (gbase[digit_of(k, P.pass)] + (uint32_t)p), k);
    }
  }
}

#ifndef ANV_SCAT_TPC
#define ANV_SCAT_TPC 1
#endif
constexpr int SCAT_TPC = ANV_SCAT_TPC;   // tiles per scatter CTA (tuning knob)
template <typename K>
__global__ void __launch_bounds__(SCAT_THREADS, ANV_SCAT_MINB) sort_scatter_kernel(const SortParams<K> P) {
  const int c = blockIdx.y;
  const ColState& S = P.state[c];
  if (S.skip[P.pass]) return;
  const int64_t n = (int64_t)S.n_valid;
  __shared__ ScatShared<K> SH;
  for (int tt = 0; tt < SCAT_TPC; ++tt) {
    const int tile = blockIdx.x * SCAT_TPC + tt;
    const int64_t t0 = (int64_t)tile * SORT_TILE;
    if (t0 >= n) return;
    const int nt = (int)min((int64_t)SORT_TILE, n - t0);
    if (nt == SORT_TILE) scatter_tile<K, true, false>(P, S, c, tile, t0, nt, SH);
    else scatter_tile<K, false, false>(P, S, c, tile, t0, nt, SH);
    __syncthreads();            // the staging area and the digit tables are reused by the next tile
  }
}

// One-sweep pass: the same stable scatter, but the tile's global digit offsets come from the column-wide digit bases
// (sort_bases_kernel) + a decoupled look-back over the preceding tiles' digit counts - one read and one write of the keys per
// pass, no tile-histogram kernel, no scan kernel.  Tile ids are tickets (arrival order), so every tile a CTA waits for has
// already started: the look-back cannot deadlock.
template <typename K>
__global__ void __launch_bounds__(SCAT_THREADS, 2) sort_onesweep_kernel(const SortParams<K> P) {
  const int c = blockIdx.y;
  const ColState& S = P.state[c];
  if (S.skip[P.pass]) return;
  __shared__ int s_tile;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&P.ticket[(size_t)c * sizeof(K) + P.pass], 1u);
  __syncthreads();
  const int tile = s_tile;
  const int64_t n = (int64_t)S.n_valid;
  const int64_t t0 = (int64_t)tile * SORT_TILE;
  if (t0 >= n) return;
  const int nt = (int)min((int64_t)SORT_TILE, n - t0);
  __shared__ ScatShared<K> SH;
  if (nt == SORT_TILE) scatter_tile<K, true, true>(P, S, c, tile, t0, nt, SH);
  else scatter_tile<K, false, true>(P, S, c, tile, t0, nt, SH);
}

// Associative combine of two ADJACENT run summaries (left, right) of sorted keys.
template <typename K> __device__ __forceinline__ void best_of(K& bk, uint32_t& bl, K k, uint32_t len) {
  if (len > bl) { bl = len; bk = k; }  // candidates arrive in ascending key order: strict > keeps the smallest key
}
template <typename K>
__device__ __forceinline__ TileSummary<K> combine(const TileSummary<K>& L, const TileSummary<K>& R) {
  if (L.n == 0) return R;
  if (R.n == 0) return L;
  TileSummary<K> o;
  const bool same = L.last_key == R.first_key;
  const bool Ls = L.prefix_len == L.n, Rs = R.prefix_len == R.n;  // the whole side is one run
  o.first_key = L.first_key; o.last_key = R.last_key; o.n = L.n + R.n;
  o.heads_inside = L.heads_inside + R.heads_inside + (same ? 0u : 1u);
  o.prefix_len = (Ls && same) ? L.n + R.prefix_len : L.prefix_len;
  o.suffix_len = (Rs && same) ? R.n + L.suffix_len : R.suffix_len;
  K bk = 0; uint32_t bl = 0;
  if (L.best_len) best_of(bk, bl, L.best_key, L.best_len);
  if (same) {
    if (!Ls && !Rs) best_of(bk, bl, L.last_key, L.suffix_len + R.prefix_len);
  } else {
    if (!Ls) best_of(bk, bl, L.last_key, L.suffix_len);
    if (!Rs) best_of(bk, bl, R.first_key, R.prefix_len);
  }
  if (R.best_len) best_of(bk, bl, R.best_key, R.best_len);
  o.best_key = bk; o.best_len = bl;
  return o;
}
template <typename K> __device__ __forceinline__ K shfl_down_key(K v, int d) {
  if (sizeof(K) == 8) return (K)__shfl_down_sync(ANV_FULL, (unsigned long long)v, d);
  return (K)__shfl_down_sync(ANV_FULL, (uint32_t)v, d);
}
template <typename K> __device__ __forceinline__ TileSummary<K> shfl_down_summary(const TileSummary<K>& s, int d) {
  TileSummary<K> r;
  r.first_key = shfl_down_key(s.first_key, d); r.last_key = shfl_down_key(s.last_key, d);
  r.best_key = shfl_down_key(s.best_key, d);
  r.n = __shfl_down_sync(ANV_FULL, s.n, d); r.prefix_len = __shfl_down_sync(ANV_FULL, s.prefix_len, d);
  r.suffix_len = __shfl_down_sync(ANV_FULL, s.suffix_len, d); r.best_len = __shfl_down_sync(ANV_FULL, s.best_len, d);
  r.heads_inside = __shfl_down_sync(ANV_FULL, s.heads_inside, d);
  return r;
}

// ---- run summary of the sorted keys --------------------------------------------------------------
// Each thread summarises 16 CONSECUTIVE sorted keys in registers (blocked 128-bit loads), the 256
// thread summaries are folded with the associative `combine` (shuffle tree per warp, then 8 warps
// sequentially): no shared-memory tile, no binary search, ~25 instructions per key.
#ifndef ANV_RUN_TPC
#define ANV_RUN_TPC 1
#endif
constexpr int RUN_TPC = ANV_RUN_TPC;     // tiles per run-summary CTA (tuning knob)
template <typename K>
__device__ __forceinline__ void run_tile_one(const SortParams<K>& P, const int c, const int tile) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ColState& S = P.state[c];
  const int64_t n = (int64_t)S.n_valid;
  const int64_t t0 = (int64_t)tile * SORT_TILE;
  TileSummary<K>& out = P.summ[(size_t)c * P.n_tiles + tile];
  if (t0 >= n) { if (tid == 0) out.n = 0; return; }
  const int nt = (int)min((int64_t)SORT_TILE, n - t0);
  const K* __restrict__ keys = (S.cur ? P.buf[1] : P.buf[0]) + (size_t)c * P.stride + t0;
  constexpr int PER = SORT_TILE / ANV_BLOCK;  // 16
  constexpr int KV = 16 / sizeof(K);
  const int m = max(0, min(PER, nt - tid * PER));
  K k[PER];
  if (m == PER) {
    const uint4* p = reinterpret_cast<const uint4*>(keys + tid * PER);
#pragma unroll
    for (int v = 0; v < PER / KV; ++v) {
      const uint4 q = __ldg(p + v);
      if (sizeof(K) == 4) { k[v * 4] = (K)q.x; k[v * 4 + 1] = (K)q.y; k[v * 4 + 2] = (K)q.z; k[v * 4 + 3] = (K)q.w; }
      else { k[v * KV] = (K)(((uint64_t)q.y << 32) | q.x); k[v * KV + (KV > 1 ? 1 : 0)] = (K)(((uint64_t)q.w << 32) | q.z); }
    }
  } else {
#pragma unroll
    for (int j = 0; j < PER; ++j) k[j] = (j < m) ? keys[tid * PER + j] : (K)0;
  }
  TileSummary<K> s;
  s.n = m; s.first_key = k[0]; s.best_key = 0; s.best_len = 0; s.heads_inside = 0; s.prefix_len = 0;
  uint32_t run = 1;
  K last = k[0];
#pragma unroll
  for (int j = 1; j < PER; ++j) {
    if (j < m) {
      if (k[j] == k[j - 1]) {
        ++run;
      } else {
        if (s.heads_inside == 0) s.prefix_len = run;
        else if (run > s.best_len) { s.best_len = run; s.best_key = k[j - 1]; }
        ++s.heads_inside;
        run = 1;
      }
      last = k[j];
    }
  }
  s.last_key = last;
  s.suffix_len = m ? run : 0;
  if (s.heads_inside == 0) s.prefix_len = m;
  // HLL++ by-product: the registers are a max over the SET of values, so hashing one key per run (plus this thread's first
  // key, whose run may have started in the previous thread - a harmless repeat) replaces the pass over all values
  extern __shared__ __align__(16) uint32_t hll_sh[];
  const int hp = P.hll_p;
  if (hp) {
    // The run heads are first COMPACTED into shared memory (a thread holds between 1 and 16 of them: hashing in place would
    // keep every warp busy for the maximum over its lanes), then hashed by full warps straight against the column's global
    // registers: 2^p words that live in L1 / L2, read before the (rare) atomicMax.  A stale cached register can only be too
    // LOW (registers never decrease), which costs a redundant atomic, never a wrong result - so there is no per-tile copy
    // of the registers to clear and merge.
    K* hk = reinterpret_cast<K*>(hll_sh);                      // [SORT_TILE] head keys
    __shared__ uint32_t s_hw[ANV_WARPS + 1];
    uint32_t hcnt = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) hcnt += (j < m && (j == 0 || k[j] != k[j - 1])) ? 1u : 0u;
    uint32_t inc = hcnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(ANV_FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_hw[warp] = inc;
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      for (int w2 = 0; w2 < ANV_WARPS; ++w2) { const uint32_t t = s_hw[w2]; s_hw[w2] = acc; acc += t; }
      s_hw[ANV_WARPS] = acc;
    }
    __syncthreads();
    uint32_t at = s_hw[warp] + inc - hcnt;
#pragma unroll
    for (int j = 0; j < PER; ++j)
      if (j < m && (j == 0 || k[j] != k[j - 1])) hk[at++] = k[j];
    __syncthreads();
    const uint32_t n_heads = s_hw[ANV_WARPS];
    const int dt = P.cols[c].dtype;
    uint32_t* G = P.hll_regs + ((size_t)c << hp);
    for (uint32_t i = tid; i < n_heads; i += ANV_BLOCK) {
      uint32_t idx, rho;
      hll_slot(spark_hash_of_key<K>(hk[i], dt), hp, idx, rho);
      if (rho > G[idx]) atomicMax(&G[idx], rho);
    }
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const TileSummary<K> right = shfl_down_summary(s, o);
    if ((lane & (2 * o - 1)) == 0) s = combine(s, right);
  }
  __shared__ TileSummary<K> ws[ANV_WARPS];
  if (lane == 0) ws[warp] = s;
  __syncthreads();
  if (tid == 0) {
    TileSummary<K> acc = ws[0];
#pragma unroll
    for (int w = 1; w < ANV_WARPS; ++w) acc = combine(acc, ws[w]);
    out = acc;
  }
}

template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK) run_tile_kernel(const SortParams<K> P) {
  for (int tt = 0; tt < RUN_TPC; ++tt) {
    const int tile = blockIdx.x * RUN_TPC + tt;
    if (tile >= P.n_tiles) return;
    run_tile_one<K>(P, blockIdx.y, tile);
    __syncthreads();           // the shared staging (head keys, warp summaries) is reused by the next tile
  }
}

// One warp per column: 32 tile summaries per step are combined with an order-preserving
// shuffle tree, the chunk results sequentially by lane 0.  Also reads the requested order
// statistics straight out of the sorted keys.
constexpr int MERGE_WARPS = 8;
template <typename K>
__global__ void __launch_bounds__(32 * MERGE_WARPS) run_merge_kernel(const SortParams<K> P, double* mode_value, int64_t* mode_rows,
                                                       int64_t* n_distinct, const int64_t* ranks, int n_ranks,
                                                       double* rank_values) {
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const ColState& S = P.state[c];
  const int64_t n = (int64_t)S.n_valid;      // sorted (nonzero) keys
  const int64_t nz = (int64_t)S.n_zero;      // the zero run that pack_kernel kept out of the sort
  const int dt = P.cols[c].dtype;
  constexpr K ZERO_KEY = (K)1 << (sizeof(K) * 8 - 1);
  const K* __restrict__ sorted = (S.cur ? P.buf[1] : P.buf[0]) + (size_t)c * P.stride;
  int64_t below = 0;                          // keys smaller than zero: the zero run occupies ranks below+1 .. below+nz
  if (nz > 0 && n_ranks > 0) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (sorted[mid] < ZERO_KEY) lo = mid + 1; else hi = mid;
    }
    below = lo;
  }
  for (int r = tid; r < n_ranks; r += 32 * MERGE_WARPS) {
    const int64_t rk = ranks[(size_t)c * n_ranks + r];
    double v = nan("");
    if (rk > 0 && rk <= n + nz) {
      K k;
      if (rk <= below) k = sorted[rk - 1];
      else if (rk <= below + nz) k = ZERO_KEY;
      else k = sorted[rk - 1 - nz];
      v = sorted_key_to_double(sizeof(K) == 8 ? (uint64_t)k : ((uint64_t)k << 32), dt);
    }
    rank_values[(size_t)c * n_ranks + r] = v;
  }
  if (P.hll_p && nz > 0 && tid == 0) {   // the zero run never reached the sort: its value hashes here
    uint32_t idx, rho;
    hll_slot(spark_hash_of_key<K>(ZERO_KEY, dt), P.hll_p, idx, rho);
    atomicMax(&P.hll_regs[((size_t)c << P.hll_p) + idx], rho);
  }
  if (S.error) {                      // a look-back gave up: make the host raise (results would be garbage)
    if (tid == 0) { mode_value[c] = nan(""); mode_rows[c] = -3; n_distinct[c] = -3; }
    return;
  }
  if (n == 0) {
    if (tid == 0) {
      mode_value[c] = nz ? sorted_key_to_double(sizeof(K) == 8 ? (uint64_t)ZERO_KEY : ((uint64_t)ZERO_KEY << 32), dt) : nan("");
      mode_rows[c] = nz;
      n_distinct[c] = nz ? 1 : 0;
    }
    return;
  }
  const TileSummary<K>* T = P.summ + (size_t)c * P.n_tiles;
  const int tiles = (int)((n + SORT_TILE - 1) / SORT_TILE);
  // every warp folds a contiguous range of tile summaries (order-preserving shuffle tree over 32 at a time), warp 0 then
  // folds the warp results in order: 8 warps instead of one walk the 24 K summaries of a 100 M-row column
  const int wid = tid >> 5;
  const int per_warp = (tiles + MERGE_WARPS - 1) / MERGE_WARPS;
  const int w_lo = wid * per_warp, w_hi = min(tiles, w_lo + per_warp);
  TileSummary<K> acc;
  acc.n = 0;
  for (int t0 = w_lo; t0 < w_hi; t0 += 32) {
    TileSummary<K> mine;
    mine.n = 0;
    if (t0 + lane < w_hi) mine = T[t0 + lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const TileSummary<K> right = shfl_down_summary(mine, o);
      if ((lane & (2 * o - 1)) == 0) mine = combine(mine, right);
    }
    if (lane == 0) acc = combine(acc, mine);
  }
  __shared__ TileSummary<K> s_acc[MERGE_WARPS];
  if (lane == 0) s_acc[wid] = acc;
  __syncthreads();
  if (tid != 0) return;
  acc = s_acc[0];
  for (int w2 = 1; w2 < MERGE_WARPS; ++w2) acc = combine(acc, s_acc[w2]);
  if (lane == 0) {
    K bk = 0; uint32_t bl = 0;
    best_of(bk, bl, acc.first_key, acc.prefix_len);
    if (acc.best_len) best_of(bk, bl, acc.best_key, acc.best_len);
    if (acc.prefix_len != acc.n) best_of(bk, bl, acc.last_key, acc.suffix_len);
    int64_t rows = bl;
    if (nz > rows || (nz == rows && ZERO_KEY < bk)) { bk = ZERO_KEY; rows = nz; }   // ties: the smaller value
    mode_value[c] = sorted_key_to_double(sizeof(K) == 8 ? (uint64_t)bk : ((uint64_t)bk << 32), dt);
    mode_rows[c] = rows;
    n_distinct[c] = (int64_t)acc.heads_inside + 1 + (nz > 0 ? 1 : 0);
  }
}

template <typename K> struct Layout {
  size_t state, buf0, buf1, tile_hist, summ, totals, ghist, gbase, ticket, status, total;
  Layout(int n_cols, int64_t n_rows) {
    const int64_t stride = (n_rows + 63) & ~(int64_t)63;
    const int64_t n_tiles = (n_rows + SORT_TILE - 1) / SORT_TILE;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
    state = take((size_t)n_cols * sizeof(ColState));
    buf0 = take((size_t)n_cols * stride * sizeof(K));
    buf1 = take((size_t)n_cols * stride * sizeof(K));
    tile_hist = take((size_t)n_cols * 256 * (n_tiles > 0 ? n_tiles : 1) * 4);
    summ = take((size_t)n_cols * (n_tiles > 0 ? n_tiles : 1) * sizeof(TileSummary<K>));
    totals = take((size_t)n_cols * 256 * 4);
    ghist = take((size_t)n_cols * sizeof(K) * 256 * 4);     // ghist, gbase, ticket, status are contiguous: one memset
    gbase = take((size_t)n_cols * sizeof(K) * 256 * 4);
    ticket = take((size_t)n_cols * sizeof(K) * 4);
    status = take((size_t)n_cols * (n_tiles > 0 ? n_tiles : 1) * 256 * 8);
    total = o + 256;
  }
};

template <typename K>
static int run_mode_distinct(const anv_column_t* cols, int n_cols, int64_t n_rows, double* mode_value, int64_t* mode_rows,
                             int64_t* n_distinct, const int64_t* ranks, int n_ranks, double* rank_values, int hll_p,
                             uint32_t* hll_regs, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  Layout<K> L(n_cols, n_rows);
  if (workspace_bytes < L.total) { set_error("anv_mode_distinct: workspace too small (%zu < %zu)", workspace_bytes, L.total); return ANV_ERR_WORKSPACE; }
  char* w = reinterpret_cast<char*>(workspace);
  SortParams<K> P{};
  P.cols = cols; P.n_cols = n_cols; P.n_rows = n_rows;
  P.stride = (n_rows + 63) & ~(int64_t)63;
  P.n_tiles = (int)((n_rows + SORT_TILE - 1) / SORT_TILE);
  if (P.n_tiles < 1) P.n_tiles = 1;
  P.buf[0] = reinterpret_cast<K*>(w + L.buf0);
  P.buf[1] = reinterpret_cast<K*>(w + L.buf1);
  P.state = reinterpret_cast<ColState*>(w + L.state);
  P.tile_hist = reinterpret_cast<uint32_t*>(w + L.tile_hist);
  P.summ = reinterpret_cast<TileSummary<K>*>(w + L.summ);
  uint32_t* totals = reinterpret_cast<uint32_t*>(w + L.totals);
  P.ghist = nullptr;                                   // set below when the one-sweep passes are selected
  P.gbase = reinterpret_cast<uint32_t*>(w + L.gbase);
  P.ticket = reinterpret_cast<uint32_t*>(w + L.ticket);
  P.status = reinterpret_cast<unsigned long long*>(w + L.status);
  P.hll_p = hll_regs ? hll_p : 0;
  P.hll_regs = hll_regs;
  if (hll_regs) ANV_CUDA(cudaMemsetAsync(hll_regs, 0, ((size_t)n_cols << hll_p) * sizeof(uint32_t), st));
  // Default: three kernels per pass (tile histogram, (digit, column)-parallel scan, stable scatter).  ANV_SORT_ONESWEEP=1 selects the
  // one-sweep passes for 32-bit keys (digit histograms in pack + decoupled look-back in the scatter: 10 instead of 14 words of
  // traffic per key).  Measured on B200 (c2, 4e8 keys): 3.06 ms per one-sweep pass against 2.19 ms for the three kernels - the
  // look-back walks ~6 predecessor tiles per digit with dependent L2 round trips while the scatter is issue-bound, not
  // HBM-bound, so removing a read of the keys buys nothing here.  Kept, tested, not the default.
  const char* os_env = getenv("ANV_SORT_ONESWEEP");
  const bool legacy = !(os_env && os_env[0] == '1') || sizeof(K) != 4;
  ANV_CUDA(cudaMemsetAsync(P.state, 0, (size_t)n_cols * sizeof(ColState), st));
  if (!legacy) {
    P.ghist = reinterpret_cast<uint32_t*>(w + L.ghist);
    ANV_CUDA(cudaMemsetAsync(w + L.ghist, 0, L.total - 256 - L.ghist, st));   // digit counts, bases, tickets, look-back status
  }
  if (n_rows > 0) {
    dim3 grid(P.n_tiles, n_cols);
    pack_kernel<K><<<dim3((P.n_tiles + PACK_TPC - 1) / PACK_TPC, n_cols), ANV_BLOCK, 0, st>>>(P);
    ANV_CUDA(cudaGetLastError());
    // 64-bit keys keep the three-kernel passes: ptxas (12.9) does not terminate on the uint64 instantiation of the one-sweep kernel
    if constexpr (sizeof(K) == 4) {
      if (!legacy) {
        sort_bases_kernel<K><<<n_cols, ANV_BLOCK, 0, st>>>(P);
        for (int pass = 0; pass < (int)sizeof(K); ++pass) {
          P.pass = pass;
          sort_onesweep_kernel<K><<<grid, SCAT_THREADS, 0, st>>>(P);
          ANV_CUDA(cudaGetLastError());
        }
      }
    }
    if (legacy || sizeof(K) != 4) {
      for (int pass = 0; pass < (int)sizeof(K); ++pass) {
        P.pass = pass;
        sort_hist_kernel<K><<<dim3((P.n_tiles + HIST_TPC - 1) / HIST_TPC, n_cols), ANV_BLOCK, 0, st>>>(P);
        sort_totals_kernel<K><<<dim3(256, n_cols), ANV_BLOCK, 0, st>>>(P, totals);
        sort_scan_kernel<K><<<dim3(256, n_cols), ANV_BLOCK, 0, st>>>(P, totals);
        sort_scatter_kernel<K><<<dim3((P.n_tiles + SCAT_TPC - 1) / SCAT_TPC, n_cols), SCAT_THREADS, 0, st>>>(P);
        ANV_CUDA(cudaGetLastError());
      }
    }
    const size_t run_smem = P.hll_p ? (size_t)SORT_TILE * sizeof(K) : 0;   // the compacted head keys
    run_tile_kernel<K><<<dim3((P.n_tiles + RUN_TPC - 1) / RUN_TPC, n_cols), ANV_BLOCK, run_smem, st>>>(P);
    ANV_CUDA(cudaGetLastError());
  }
  run_merge_kernel<K><<<n_cols, 32 * MERGE_WARPS, 0, st>>>(P, mode_value, mode_rows, n_distinct, ranks, n_ranks, rank_values);
  ANV_CUDA(cudaGetLastError());
  return ANV_OK;
}


// =====================================================================================================================
// Partition + count path for 32-bit keys (F32 / I32 columns): what mode / distinct / percentiles need is the multiset of
// keys in key ORDER, not a sorted array - so the columns are not sorted at all:
//   sample     32 * P keys per column (stratified row positions), sorted with the LSD kernels above (tiny);
//   split      P - 1 splitters = every 32nd sample key, plus the key of zero as a forced splitter, and a 4096-cell
//              lookup table over the top 12 key bits that narrows the splitter search to a few steps;
//   partition  ONE read of the raw column: every key finds its bucket (lower_bound over the splitters).  A key EQUAL to a
//              splitter is only counted (warp-aggregated atomics) - any value frequent enough to unbalance a bucket is a
//              splitter with overwhelming probability (it occupies >= 32 sample slots), so heavy hitters, exact zeros and
//              discrete-valued columns never reach the key buffer; the other keys are appended to their bucket's slab
//              (capacity 4x the mean bucket; the open 32-byte sectors of all buckets stay in L2 until they are full);
//   cum        prefix sums over the interleaved (bucket, splitter) counts: total order of the column => every requested
//              rank resolves to a splitter value directly or to (bucket, local rank);
//   count      one CTA per bucket: shared-memory hash table key -> multiplicity (buckets larger than the table are swept
//              several times, each sweep taking one hash class), distinct count and longest run per bucket, in-bucket radix
//              select for the few buckets that hold a requested rank.
// HBM traffic: one read of the column + ~1 write and 1 read of the keys that are not splitters (vs ~14 words per key for
// the LSD path); no ranking, no stable scatter.  Counting is integer everywhere => deterministic results (the slab order is
// not, and does not matter).  A bucket that overflows its slab or a hash table that fills up raises a per-column flag; the
// host then redoes that column on the LSD path.
constexpr uint32_t PC_ZERO_KEY = 0x80000000u;    // key of +-0.0 / integer 0: always a splitter, doubles as EMPTY in the hash table
constexpr int PC_OVERSAMPLE = 32;
constexpr int PC_LUT_BITS = 12;
constexpr int PC_LUT_CELLS = 1 << PC_LUT_BITS;
constexpr int PC_SLOTS_LOG = 13;
constexpr int PC_SLOTS = 1 << PC_SLOTS_LOG;      // hash table slots per CTA (64 KB: keys + counts)
constexpr int PC_SWEEP_KEYS = 3072;              // keys one sweep of the table is sized for
constexpr int PC_TILES_PER_CTA = 32;             // partition kernel: 32 x 4096 rows per CTA (amortises the splitter load)
constexpr int PC_MAX_RANKS = 16;

struct PcCol {                     // per column, in the workspace (zeroed per call)
  unsigned long long n_valid;      // non-null values
  unsigned long long distinct;     // sum of the buckets' distinct counts + splitters that occur
  unsigned long long best;         // max over (multiplicity << 32 | ~key): the mode, ties -> smallest key
  int overflow;
  int n_queries;                   // ranks that fall inside a bucket
  int q_bucket[PC_MAX_RANKS];
  uint32_t q_local[PC_MAX_RANKS];  // 1-based rank inside the bucket
  int q_slot[PC_MAX_RANKS];        // index into rank_values
};

struct PcParams {
  const anv_column_t* cols;
  int n_cols;
  int64_t n_rows;
  int P, NS, NB;                   // NS = P splitters (P - 1 from the sample + zero), NB = NS + 1 buckets
  uint32_t cap;                    // slab capacity per bucket (keys)
  int64_t m;                       // sample slots per column
  uint32_t* split;                 // [n_cols][NS]
  uint16_t* lut;                   // [n_cols][PC_LUT_CELLS + 1]
  uint32_t* cursor;                // [n_cols][NB]  keys appended to each bucket
  uint32_t* cnt_eq;                // [n_cols][NS]  keys equal to each splitter
  uint32_t* cum;                   // [n_cols][2 * NB]  inclusive prefix over lt_0, eq_0, lt_1, eq_1, ...
  uint32_t* slab;                  // [n_cols][NB][cap]
  PcCol* st;
};

template <typename T> __device__ __forceinline__ T load_elem(const void* base, int64_t row) {
  return reinterpret_cast<const T*>(base)[row];
}

// ---- sample: m stratified row positions per column -> keys (nulls and zeros dropped), compacted into the LSD buffers ----
__global__ void __launch_bounds__(ANV_BLOCK) pc_sample_kernel(const SortParams<uint32_t> S, const int64_t n_rows, const int64_t m) {
  const int c = blockIdx.y;
  const anv_column_t col = S.cols[c];
  const int64_t i = (int64_t)blockIdx.x * ANV_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool ok = false;
  uint32_t key = 0;
  if (i < m) {
    const int64_t stride = n_rows / m > 0 ? n_rows / m : 1;
    uint32_t h = (uint32_t)i * 0x9E3779B1u + (uint32_t)c * 0x85EBCA6Bu;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    const int64_t row = i * stride + (int64_t)(h % (uint32_t)stride);
    if (row < n_rows) {
      ok = col.validity ? ((col.validity[row >> 5] >> (row & 31)) & 1u) : true;
      if (ok) {
        key = (col.dtype == ANV_F32) ? make_key<uint32_t, float>(load_elem<float>(col.data, row))
                                     : make_key<uint32_t, int32_t>(load_elem<int32_t>(col.data, row));
        ok = key != PC_ZERO_KEY;
      }
    }
  }
  __shared__ uint32_t s_w[ANV_WARPS];
  __shared__ unsigned long long s_base;
  const uint32_t bal = __ballot_sync(ANV_FULL, ok);
  if (lane == 0) s_w[warp] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int w = 0; w < ANV_WARPS; ++w) { const uint32_t t = s_w[w]; s_w[w] = acc; acc += t; }
    s_base = acc ? atomicAdd(&S.state[c].n_valid, (unsigned long long)acc) : 0ull;
  }
  __syncthreads();
  if (ok) S.buf[0][(size_t)c * S.stride + s_base + s_w[warp] + __popc(bal & ((1u << lane) - 1u))] = key;
}

// ---- split: splitters + lookup table of one column ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) pc_split_kernel(const SortParams<uint32_t> S, const PcParams P) {
  const int c = blockIdx.x, tid = threadIdx.x;
  const ColState& st = S.state[c];
  const uint32_t* __restrict__ sorted = (st.cur ? S.buf[1] : S.buf[0]) + (size_t)c * S.stride;
  const int64_t ms = (int64_t)st.n_valid;          // sample keys that survived (non-null, nonzero)
  const int ns = P.P - 1;                          // splitters taken from the sample
  auto sample_split = [&](int j) -> uint32_t {     // non-decreasing in j
    if (ms == 0) return PC_ZERO_KEY;
    int64_t idx = ((int64_t)(j + 1) * ms) / P.P;
    if (idx >= ms) idx = ms - 1;
    return sorted[idx];
  };
  int lo = 0, hi = ns;                             // z = number of sample splitters below the key of zero
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (sample_split(mid) < PC_ZERO_KEY) lo = mid + 1; else hi = mid; }
  const int z = lo;
  uint32_t* out = P.split + (size_t)c * P.NS;
  for (int j = tid; j < ns; j += 256) { const uint32_t v = sample_split(j); out[j < z ? j : j + 1] = v; }
  if (tid == 0) out[z] = PC_ZERO_KEY;
  __syncthreads();
  uint16_t* lut = P.lut + (size_t)c * (PC_LUT_CELLS + 1);
  for (int q = tid; q <= PC_LUT_CELLS; q += 256) {
    int a = 0, b = P.NS;
    if (q == PC_LUT_CELLS) { a = P.NS; }
    else {
      const uint32_t v = (uint32_t)q << (32 - PC_LUT_BITS);
      while (a < b) { const int mid = (a + b) >> 1; if (out[mid] < v) a = mid + 1; else b = mid; }
    }
    lut[q] = (uint16_t)a;
  }
}

// ---- partition ----------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void pc_part_rows(const PcParams& P, const anv_column_t& col, const int c, const uint32_t* sS, const uint16_t* sL,
                                             const int64_t r0, const int n_tile, unsigned long long& valid_acc) {
  constexpr int VEC = 4, PER = SORT_TILE / ANV_BLOCK;
  const int tid = threadIdx.x, lane = tid & 31;
  const T* __restrict__ data = reinterpret_cast<const T*>(col.data) + r0;
  const uint32_t* __restrict__ vbits = col.validity;
  const int row0 = tid * PER;
  uint32_t keys[PER];
  uint32_t okmask = 0;
  if (row0 + PER <= n_tile) {
    uint32_t vb = 0xFFFFu;
    if (vbits) { const int64_t g = r0 + row0; vb = (__ldg(vbits + (g >> 5)) >> (g & 31)) & 0xFFFFu; }
    okmask = vb;
    const uint4* p = reinterpret_cast<const uint4*>(data + row0);
#pragma unroll
    for (int v = 0; v < PER / VEC; ++v) {
      const uint4 q = ldg_stream(p + v);
      T e[VEC];
      unpack<T>(q, e);
#pragma unroll
      for (int i = 0; i < VEC; ++i) keys[v * VEC + i] = make_key<uint32_t, T>(e[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int row = row0 + i;
      bool ok = row < n_tile;
      keys[i] = 0;
      if (ok) {
        if (vbits) { const int64_t g = r0 + row; ok = (vbits[g >> 5] >> (g & 31)) & 1u; }
        keys[i] = make_key<uint32_t, T>(data[row]);
      }
      okmask |= ok ? (1u << i) : 0u;
    }
  }
  valid_acc += __popc(okmask);
  uint32_t* __restrict__ cursor = P.cursor + (size_t)c * P.NB;
  uint32_t* __restrict__ cnt_eq = P.cnt_eq + (size_t)c * P.NS;
  uint32_t* __restrict__ slab = P.slab + (size_t)c * P.NB * P.cap;
  bool over = false;
#pragma unroll 4
  for (int i = 0; i < PER; ++i) {
    const bool ok = (okmask >> i) & 1u;
    const uint32_t k = keys[i];
    int lo = sL[k >> (32 - PC_LUT_BITS)], hi = sL[(k >> (32 - PC_LUT_BITS)) + 1];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sS[mid] < k) lo = mid + 1; else hi = mid;
    }
    const bool eq = ok && lo < P.NS && sS[lo] == k;
    const uint32_t em = __ballot_sync(ANV_FULL, eq);
    if (eq) {
      const uint32_t peers = __match_any_sync(em, lo);
      if (lane == __ffs(peers) - 1) atomicAdd(&cnt_eq[lo], (uint32_t)__popc(peers));
    } else if (ok) {
      const uint32_t pos = atomicAdd(&cursor[lo], 1u);
      if (pos < P.cap) slab[(size_t)lo * P.cap + pos] = k;
      else over = true;
    }
  }
  if (over) P.st[c].overflow = 1;
}

__global__ void __launch_bounds__(ANV_BLOCK) pc_partition_kernel(const PcParams P) {
  extern __shared__ __align__(16) uint32_t pc_sh[];
  const int c = blockIdx.y, tid = threadIdx.x;
  const anv_column_t col = P.cols[c];
  uint32_t* sS = pc_sh;                                               // [NS]
  uint16_t* sL = reinterpret_cast<uint16_t*>(pc_sh + ((P.NS + 3) & ~3));   // [PC_LUT_CELLS + 1]
  {
    const uint32_t* gS = P.split + (size_t)c * P.NS;
    for (int i = tid; i < P.NS; i += ANV_BLOCK) sS[i] = gS[i];
    const uint16_t* gL = P.lut + (size_t)c * (PC_LUT_CELLS + 1);
    for (int i = tid; i <= PC_LUT_CELLS; i += ANV_BLOCK) sL[i] = gL[i];
  }
  __syncthreads();
  unsigned long long valid_acc = 0;
  const int64_t first = (int64_t)blockIdx.x * PC_TILES_PER_CTA;
  for (int t = 0; t < PC_TILES_PER_CTA; ++t) {
    const int64_t r0 = (first + t) * SORT_TILE;
    if (r0 >= P.n_rows) break;
    const int n_tile = (int)min((int64_t)SORT_TILE, P.n_rows - r0);
    if (col.dtype == ANV_F32) pc_part_rows<float>(P, col, c, sS, sL, r0, n_tile, valid_acc);
    else pc_part_rows<int32_t>(P, col, c, sS, sL, r0, n_tile, valid_acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) valid_acc += __shfl_down_sync(ANV_FULL, valid_acc, o);
  if ((tid & 31) == 0 && valid_acc) atomicAdd(&P.st[c].n_valid, valid_acc);
}

// ---- cum: total order of the column from the interleaved counts; rank queries; splitter contributions ---------------------
__global__ void __launch_bounds__(1024) pc_cum_kernel(const PcParams P, const int64_t* ranks, const int n_ranks, double* rank_values) {
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  PcCol& st = P.st[c];
  const uint32_t* cursor = P.cursor + (size_t)c * P.NB;
  const uint32_t* cnt_eq = P.cnt_eq + (size_t)c * P.NS;
  const uint32_t* S = P.split + (size_t)c * P.NS;
  uint32_t* cum = P.cum + (size_t)c * 2 * P.NB;
  const int n_ent = 2 * P.NB - 1;                                     // lt_0, eq_0, lt_1, ..., eq_{NS-1}, lt_NS
  const int per = (n_ent + 1023) / 1024;
  const int e0 = tid * per, e1 = min(e0 + per, n_ent);
  auto entry = [&](int e) -> uint32_t { return (e & 1) ? cnt_eq[e >> 1] : min(cursor[e >> 1], P.cap); };
  unsigned long long run = 0, eq_distinct = 0, best = 0;
  bool over = false;
  for (int e = e0; e < e1; ++e) {
    run += entry(e);
    if (e & 1) {
      const uint32_t n = cnt_eq[e >> 1];
      if (n) { ++eq_distinct; const unsigned long long cand = ((unsigned long long)n << 32) | (uint32_t)~S[e >> 1]; if (cand > best) best = cand; }
    } else if (cursor[e >> 1] > P.cap) over = true;
  }
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long wbest[32], wdist[32];
  unsigned long long inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(ANV_FULL, inc, o); if (lane >= o) inc += t; }
  unsigned long long rb = best, rd = eq_distinct;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ob = __shfl_down_sync(ANV_FULL, rb, o); if (ob > rb) rb = ob;
    rd += __shfl_down_sync(ANV_FULL, rd, o);
  }
  if (lane == 31) wsum[warp] = inc;
  if (lane == 0) { wbest[warp] = rb; wdist[warp] = rd; }
  if (over) st.overflow = 1;
  __syncthreads();
  if (warp == 0) {
    const unsigned long long w = wsum[lane];
    unsigned long long wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(ANV_FULL, wi, o); if (lane >= o) wi += t; }
    wsum[lane] = wi - w;
    unsigned long long b2 = wbest[lane], d2 = wdist[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ob = __shfl_down_sync(ANV_FULL, b2, o); if (ob > b2) b2 = ob;
      d2 += __shfl_down_sync(ANV_FULL, d2, o);
    }
    if (lane == 0) { if (b2) atomicMax(&st.best, b2); if (d2) atomicAdd(&st.distinct, d2); }
  }
  __syncthreads();
  unsigned long long acc = wsum[warp] + inc - run;
  for (int e = e0; e < e1; ++e) { acc += entry(e); cum[e] = (uint32_t)acc; }
  __syncthreads();
  // rank queries: first entry whose inclusive prefix reaches the rank
  if (tid < n_ranks) {
    const int64_t rk = ranks[(size_t)c * n_ranks + tid];
    const unsigned long long total = n_ent > 0 ? cum[n_ent - 1] : 0;
    double v = nan("");
    if (rk > 0 && (unsigned long long)rk <= total) {
      int lo = 0, hi = n_ent - 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if ((unsigned long long)cum[mid] < (unsigned long long)rk) lo = mid + 1; else hi = mid; }
      if (lo & 1) {
        v = sorted_key_to_double((uint64_t)S[lo >> 1] << 32, P.cols[c].dtype);
      } else {
        const int q = atomicAdd(&st.n_queries, 1);
        st.q_bucket[q] = lo >> 1;
        st.q_local[q] = (uint32_t)(rk - (lo ? cum[lo - 1] : 0));
        st.q_slot[q] = tid;
      }
    }
    rank_values[(size_t)c * n_ranks + tid] = v;
  }
}

// ---- count: one CTA per bucket ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ANV_BLOCK) pc_count_kernel(const PcParams P, const int n_ranks, double* rank_values) {
  const int c = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
  PcCol& st = P.st[c];
  const uint32_t n = min(P.cursor[(size_t)c * P.NB + b], P.cap);
  if (n == 0) return;                                   // (uniform: the cursors are final when this kernel starts)
  extern __shared__ __align__(16) uint32_t pc_tab[];
  uint32_t* tkey = pc_tab;
  uint32_t* tcnt = pc_tab + PC_SLOTS;
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_sel[2];
  const uint32_t* __restrict__ keys = P.slab + ((size_t)c * P.NB + b) * P.cap;
  const uint32_t sweeps = (n + PC_SWEEP_KEYS - 1) / PC_SWEEP_KEYS;
  unsigned long long best = 0;
  uint32_t distinct = 0;
  bool full = false;
  for (uint32_t sw = 0; sw < sweeps; ++sw) {
    for (int i = tid; i < PC_SLOTS; i += ANV_BLOCK) { tkey[i] = PC_ZERO_KEY; tcnt[i] = 0; }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += ANV_BLOCK) {
      const uint32_t k = keys[i];
      if (sweeps > 1 && ((k * 0x85EBCA6Bu) >> 12) % sweeps != sw) continue;
      uint32_t slot = (k * 0x9E3779B1u) >> (32 - PC_SLOTS_LOG);
      int probes = 0;
      while (true) {
        const uint32_t prev = atomicCAS(&tkey[slot], PC_ZERO_KEY, k);
        if (prev == PC_ZERO_KEY || prev == k) { atomicAdd(&tcnt[slot], 1u); break; }
        slot = (slot + 1) & (PC_SLOTS - 1);
        if (++probes >= PC_SLOTS) { full = true; break; }
      }
    }
    __syncthreads();
    for (int i = tid; i < PC_SLOTS; i += ANV_BLOCK) {
      const uint32_t cnt = tcnt[i];
      if (cnt) {
        ++distinct;
        const unsigned long long cand = ((unsigned long long)cnt << 32) | (uint32_t)~tkey[i];
        if (cand > best) best = cand;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ob = __shfl_down_sync(ANV_FULL, best, o); if (ob > best) best = ob;
    distinct += __shfl_down_sync(ANV_FULL, distinct, o);
  }
  if ((tid & 31) == 0) {
    if (best) atomicMax(&st.best, best);
    if (distinct) atomicAdd(&st.distinct, (unsigned long long)distinct);
  }
  if (full) st.overflow = 1;
  // requested ranks inside this bucket: 4-pass radix select over the bucket's keys (L2-resident)
  const int nq = st.n_queries;
  for (int q = 0; q < nq; ++q) {
    if (st.q_bucket[q] != b) continue;                     // uniform across the CTA
    uint32_t prefix = 0, r = st.q_local[q];
    for (int pass = 3; pass >= 0; --pass) {
      s_hist[tid] = 0;
      __syncthreads();
      const int sh = pass * 8;
      for (uint32_t i = tid; i < n; i += ANV_BLOCK) {
        const uint32_t k = keys[i];
        if (pass == 3 || ((k ^ prefix) >> (sh + 8)) == 0) atomicAdd(&s_hist[(k >> sh) & 0xFFu], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t acc = 0, d = 0;
        for (; d < 255; ++d) { if (acc + s_hist[d] >= r) break; acc += s_hist[d]; }
        s_sel[0] = d; s_sel[1] = r - acc;
      }
      __syncthreads();
      prefix |= s_sel[0] << sh;
      r = s_sel[1];
      __syncthreads();
    }
    if (tid == 0) rank_values[(size_t)c * n_ranks + st.q_slot[q]] = sorted_key_to_double((uint64_t)prefix << 32, P.cols[c].dtype);
  }
}

__global__ void pc_final_kernel(const PcParams P, double* mode_value, int64_t* mode_rows, int64_t* n_distinct) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_cols) return;
  const PcCol& st = P.st[c];
  if (st.overflow) { mode_value[c] = nan(""); mode_rows[c] = -2; n_distinct[c] = -2; return; }   // redo on the LSD path
  if (st.n_valid == 0 || st.best == 0) { mode_value[c] = nan(""); mode_rows[c] = 0; n_distinct[c] = 0; return; }
  const uint32_t key = ~(uint32_t)(st.best & 0xFFFFFFFFull);
  mode_value[c] = sorted_key_to_double((uint64_t)key << 32, P.cols[c].dtype);
  mode_rows[c] = (int64_t)(st.best >> 32);
  n_distinct[c] = (int64_t)st.distinct;
}

struct PcLayout {
  int P, NS, NB;
  uint32_t cap;
  int64_t m;
  size_t lsd, split, lut, cursor, cnt_eq, cum, st, slab, total;
  PcLayout(int n_cols, int64_t n_rows) {
    int p = 256;
    while (p < 16384 && (int64_t)p * 3072 < n_rows) p <<= 1;
    P = p; NS = p; NB = p + 1;
    m = (int64_t)PC_OVERSAMPLE * p;
    if (m > n_rows) m = n_rows > 0 ? n_rows : 1;
    const int64_t mean = n_rows / p + 1;
    cap = (uint32_t)(((4 * mean + 1024) + 7) & ~(int64_t)7);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
    lsd = take(Layout<uint32_t>(n_cols, m).total);
    split = take((size_t)n_cols * NS * 4);
    lut = take((size_t)n_cols * (PC_LUT_CELLS + 1) * 2);
    cursor = take((size_t)n_cols * NB * 4);
    cnt_eq = take((size_t)n_cols * NS * 4);
    cum = take((size_t)n_cols * 2 * NB * 4);
    st = take((size_t)n_cols * sizeof(PcCol));
    slab = take((size_t)n_cols * NB * cap * 4);
    total = o + 256;
  }
};

static int run_partition_count(const anv_column_t* cols, int n_cols, int64_t n_rows, double* mode_value, int64_t* mode_rows,
                               int64_t* n_distinct, const int64_t* ranks, int n_ranks, double* rank_values, void* workspace,
                               size_t workspace_bytes, cudaStream_t st) {
  PcLayout L(n_cols, n_rows);
  if (workspace_bytes < L.total) { set_error("anv_mode_distinct_partition: workspace too small (%zu < %zu)", workspace_bytes, L.total); return ANV_ERR_WORKSPACE; }
  char* w = reinterpret_cast<char*>(workspace);
  // --- the sample sort reuses the LSD machinery on m keys per column ---
  Layout<uint32_t> LS(n_cols, L.m);
  char* ws = w + L.lsd;
  SortParams<uint32_t> S{};
  S.cols = cols; S.n_cols = n_cols; S.n_rows = L.m;
  S.stride = (L.m + 63) & ~(int64_t)63;
  S.n_tiles = (int)((L.m + SORT_TILE - 1) / SORT_TILE);
  if (S.n_tiles < 1) S.n_tiles = 1;
  S.buf[0] = reinterpret_cast<uint32_t*>(ws + LS.buf0);
  S.buf[1] = reinterpret_cast<uint32_t*>(ws + LS.buf1);
  S.state = reinterpret_cast<ColState*>(ws + LS.state);
  S.tile_hist = reinterpret_cast<uint32_t*>(ws + LS.tile_hist);
  S.summ = reinterpret_cast<TileSummary<uint32_t>*>(ws + LS.summ);
  uint32_t* totals = reinterpret_cast<uint32_t*>(ws + LS.totals);
  PcParams P{};
  P.cols = cols; P.n_cols = n_cols; P.n_rows = n_rows; P.P = L.P; P.NS = L.NS; P.NB = L.NB; P.cap = L.cap; P.m = L.m;
  P.split = reinterpret_cast<uint32_t*>(w + L.split);
  P.lut = reinterpret_cast<uint16_t*>(w + L.lut);
  P.cursor = reinterpret_cast<uint32_t*>(w + L.cursor);
  P.cnt_eq = reinterpret_cast<uint32_t*>(w + L.cnt_eq);
  P.cum = reinterpret_cast<uint32_t*>(w + L.cum);
  P.st = reinterpret_cast<PcCol*>(w + L.st);
  P.slab = reinterpret_cast<uint32_t*>(w + L.slab);
  ANV_CUDA(cudaMemsetAsync(S.state, 0, (size_t)n_cols * sizeof(ColState), st));
  ANV_CUDA(cudaMemsetAsync(w + L.cursor, 0, L.st + (size_t)n_cols * sizeof(PcCol) - L.cursor, st));   // cursor, cnt_eq, cum, st
  {
    dim3 grid((unsigned)((L.m + ANV_BLOCK - 1) / ANV_BLOCK), n_cols);
    pc_sample_kernel<<<grid, ANV_BLOCK, 0, st>>>(S, n_rows, L.m);
    ANV_CUDA(cudaGetLastError());
    dim3 tg(S.n_tiles, n_cols);
    for (int pass = 0; pass < 4; ++pass) {
      S.pass = pass;
      sort_hist_kernel<uint32_t><<<dim3((S.n_tiles + HIST_TPC - 1) / HIST_TPC, n_cols), ANV_BLOCK, 0, st>>>(S);
      sort_totals_kernel<uint32_t><<<dim3(256, n_cols), ANV_BLOCK, 0, st>>>(S, totals);
      sort_scan_kernel<uint32_t><<<dim3(256, n_cols), ANV_BLOCK, 0, st>>>(S, totals);
      sort_scatter_kernel<uint32_t><<<dim3((S.n_tiles + SCAT_TPC - 1) / SCAT_TPC, n_cols), SCAT_THREADS, 0, st>>>(S);
      ANV_CUDA(cudaGetLastError());
    }
    pc_split_kernel<<<n_cols, 256, 0, st>>>(S, P);
    ANV_CUDA(cudaGetLastError());
  }
  {
    const size_t smem = (size_t)((L.NS + 3) & ~3) * 4 + (size_t)(PC_LUT_CELLS + 1) * 2 + 16;
    static bool attr_done = false;
    if (!attr_done) {
      ANV_CUDA(cudaFuncSetAttribute(pc_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      ANV_CUDA(cudaFuncSetAttribute(pc_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SLOTS * 8));
      attr_done = true;
    }
    const int64_t tiles = (n_rows + SORT_TILE - 1) / SORT_TILE;
    dim3 grid((unsigned)((tiles + PC_TILES_PER_CTA - 1) / PC_TILES_PER_CTA), n_cols);
    if (tiles > 0) pc_partition_kernel<<<grid, ANV_BLOCK, smem, st>>>(P);
    ANV_CUDA(cudaGetLastError());
  }
  pc_cum_kernel<<<n_cols, 1024, 0, st>>>(P, ranks, n_ranks, rank_values);
  ANV_CUDA(cudaGetLastError());
  {
    dim3 grid(L.NB, n_cols);
    pc_count_kernel<<<grid, ANV_BLOCK, PC_SLOTS * 8, st>>>(P, n_ranks, rank_values);
    ANV_CUDA(cudaGetLastError());
  }
  pc_final_kernel<<<(n_cols + 127) / 128, 128, 0, st>>>(P, mode_value, mode_rows, n_distinct);
  ANV_CUDA(cudaGetLastError());
  return ANV_OK;
}

}  // namespace anv

using namespace anv;

extern "C" size_t anv_mode_distinct_workspace_bytes(int n_cols, int64_t n_rows, int key_bits) {
  if (n_cols <= 0 || n_rows < 0) return 256;
  return key_bits == 32 ? Layout<uint32_t>(n_cols, n_rows).total : Layout<uint64_t>(n_cols, n_rows).total;
}

extern "C" int anv_mode_distinct_hll(const anv_column_t* cols, int n_cols, int64_t n_rows, int key_bits, double* mode_value,
                                     int64_t* mode_rows, int64_t* n_distinct, const int64_t* ranks, int n_ranks,
                                     double* rank_values, int hll_p, uint32_t* hll_regs, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (n_ranks < 0 || (n_ranks > 0 && (!ranks || !rank_values))) { set_error("anv_mode_distinct: bad ranks arguments"); return ANV_ERR_INVALID; }
  if (n_cols < 0 || n_rows < 0 || (key_bits != 32 && key_bits != 64)) { set_error("anv_mode_distinct: bad arguments"); return ANV_ERR_INVALID; }
  if (hll_regs && (hll_p < 4 || hll_p > 12)) { set_error("anv_mode_distinct_hll: 4 <= hll_p <= 12"); return ANV_ERR_INVALID; }
  if (n_cols == 0) return ANV_OK;
  if (n_cols > 65535) { set_error("n_cols > 65535"); return ANV_ERR_UNSUPPORTED; }
  if (n_rows >= ((int64_t)1 << 32)) { set_error("anv_mode_distinct: n_rows >= 2^32 per call is not supported"); return ANV_ERR_UNSUPPORTED; }
  if (!cols || !mode_value || !mode_rows || !n_distinct || !workspace) { set_error("anv_mode_distinct: NULL argument"); return ANV_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (key_bits == 32) return run_mode_distinct<uint32_t>(cols, n_cols, n_rows, mode_value, mode_rows, n_distinct, ranks, n_ranks, rank_values, hll_p, hll_regs, workspace, workspace_bytes, st);
  return run_mode_distinct<uint64_t>(cols, n_cols, n_rows, mode_value, mode_rows, n_distinct, ranks, n_ranks, rank_values, hll_p, hll_regs, workspace, workspace_bytes, st);
}

extern "C" int anv_mode_distinct(const anv_column_t* cols, int n_cols, int64_t n_rows, int key_bits, double* mode_value,
                                 int64_t* mode_rows, int64_t* n_distinct, const int64_t* ranks, int n_ranks,
                                 double* rank_values, void* workspace, size_t workspace_bytes, void* stream) {
  return anv_mode_distinct_hll(cols, n_cols, n_rows, key_bits, mode_value, mode_rows, n_distinct, ranks, n_ranks, rank_values, 0,
                               nullptr, workspace, workspace_bytes, stream);
}

extern "C" size_t anv_mode_distinct_partition_workspace_bytes(int n_cols, int64_t n_rows) {
  if (n_cols <= 0 || n_rows < 0) return 256;
  return PcLayout(n_cols, n_rows).total;
}

extern "C" int anv_mode_distinct_partition(const anv_column_t* cols, int n_cols, int64_t n_rows, double* mode_value,
                                           int64_t* mode_rows, int64_t* n_distinct, const int64_t* ranks, int n_ranks,
                                           double* rank_values, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_ranks < 0 || n_ranks > PC_MAX_RANKS || (n_ranks > 0 && (!ranks || !rank_values))) { set_error("anv_mode_distinct_partition: bad ranks arguments (n_ranks <= 16)"); return ANV_ERR_INVALID; }
  if (n_cols < 0 || n_rows < 0) { set_error("anv_mode_distinct_partition: bad arguments"); return ANV_ERR_INVALID; }
  if (n_cols == 0) return ANV_OK;
  if (n_cols > 65535) { set_error("n_cols > 65535"); return ANV_ERR_UNSUPPORTED; }
  if (n_rows >= ((int64_t)1 << 32)) { set_error("anv_mode_distinct_partition: n_rows >= 2^32 per call is not supported"); return ANV_ERR_UNSUPPORTED; }
  if (!cols || !mode_value || !mode_rows || !n_distinct || !workspace) { set_error("anv_mode_distinct_partition: NULL argument"); return ANV_ERR_INVALID; }
  return run_partition_count(cols, n_cols, n_rows, mode_value, mode_rows, n_distinct, ranks, n_ranks, rank_values, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}
