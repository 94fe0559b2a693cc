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
//   pack     values -> keys, nulls dropped (warp-aggregated unordered compaction: the order
//            before a sort is irrelevant), n_valid counted on the device;
//   8-bit LSD passes: tile histogram -> per-column exclusive scan -> stable scatter
//            (per-warp match_any ranking); a pass whose digit is constant is skipped
//            (device-side decision, no host sync) - float data rarely needs all bytes;
//   runs     per-tile run summary of the sorted keys (head count, open prefix / suffix run,
//            longest closed run) merged sequentially per column.
// Counting is integer everywhere => deterministic.  Ties for the mode resolve to the
// smallest value (the reference's choice is arbitrary, stats_generator.py:358).
#include "common.cuh"

namespace anv {

constexpr int SORT_TILE = 4096;  // keys per CTA (8 warps x 16 rounds x 32 lanes)
constexpr int SORT_ROUNDS = SORT_TILE / ANV_BLOCK;

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

struct ColState {               // one per column, in the workspace
  unsigned long long n_valid;   // filled by pack_kernel
  int cur;                      // which ping-pong buffer holds the current order
  int src[8];                   // per pass: source buffer
  int skip[8];                  // per pass: digit constant -> no scatter
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
  unsigned long long* digit_total;  // [n_cols][256]
  TileSummary<K>* summ;         // [n_cols][n_tiles]
  int pass;
};

// ---- pack: values -> keys, nulls dropped ------------------------------------------------------
template <typename K, typename T>
__device__ __forceinline__ void pack_column(const SortParams<K>& P, const anv_column_t& col, int c) {
  const T* __restrict__ data = reinterpret_cast<const T*>(col.data);
  const uint32_t* __restrict__ vbits = col.validity;
  K* __restrict__ out = P.buf[0] + (size_t)c * P.stride;
  unsigned long long* counter = &P.state[c].n_valid;
  const int lane = threadIdx.x & 31;
  for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(int64_t)31;
       base < P.n_rows; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = base + lane;
    bool ok = i < P.n_rows;
    T x = ok ? data[i] : (T)0;
    if (ok && vbits) ok = (__ldg(vbits + (i >> 5)) >> (i & 31)) & 1u;
    const uint32_t m = __ballot_sync(ANV_FULL, ok);
    if (!m) continue;
    unsigned long long pos = 0;
    if (lane == 0) pos = atomicAdd(counter, (unsigned long long)__popc(m));
    pos = __shfl_sync(ANV_FULL, pos, 0);
    if (ok) out[pos + __popc(m & ((1u << lane) - 1u))] = make_key<K, T>(x);
  }
}

template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK) pack_kernel(const SortParams<K> P) {
  const int c = blockIdx.y;
  const anv_column_t col = P.cols[c];
  switch (col.dtype) {
    case ANV_F32: pack_column<K, float>(P, col, c); break;
    case ANV_I32: pack_column<K, int32_t>(P, col, c); break;
    case ANV_F64: if (sizeof(K) == 8) pack_column<K, double>(P, col, c); break;
    case ANV_I64: if (sizeof(K) == 8) pack_column<K, int64_t>(P, col, c); break;
    default: break;
  }
}

template <typename K> __device__ __forceinline__ uint32_t digit_of(K k, int pass) {
  return (uint32_t)(k >> (pass * 8)) & 0xFFu;
}

// ---- pass step 1: per-tile digit histogram ------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK) sort_hist_kernel(const SortParams<K> P) {
  const int c = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const ColState& S = P.state[c];
  const int64_t n = (int64_t)S.n_valid;
  const int64_t t0 = (int64_t)tile * SORT_TILE;
  __shared__ uint32_t h[256];
  h[tid] = 0;
  __syncthreads();
  if (t0 < n) {
    const K* __restrict__ keys = P.buf[S.cur] + (size_t)c * P.stride;
#pragma unroll 4
    for (int r = 0; r < SORT_ROUNDS; ++r) {
      const int64_t i = t0 + r * ANV_BLOCK + tid;
      const bool ok = i < n;
      const uint32_t act = __ballot_sync(ANV_FULL, ok);
      if (ok) {
        const uint32_t d = digit_of(keys[i], P.pass);
        const uint32_t m = __match_any_sync(act, d);
        if (lane == __ffs(m) - 1) atomicAdd(&h[d], (uint32_t)__popc(m));
      }
    }
  }
  __syncthreads();
  const uint32_t v = h[tid];
  P.tile_hist[((size_t)c * 256 + tid) * P.n_tiles + tile] = v;
  if (v) atomicAdd(&P.digit_total[(size_t)c * 256 + tid], (unsigned long long)v);
}

// ---- pass step 2: per-column exclusive scan of [256][n_tiles] + skip decision ----------------------
template <typename K>
__global__ void __launch_bounds__(1024) sort_scan_kernel(const SortParams<K> P) {
  const int c = blockIdx.x, tid = threadIdx.x;
  ColState& S = P.state[c];
  __shared__ int s_skip;
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t s_carry;
  if (tid == 0) { s_skip = 0; s_carry = 0; }
  __syncthreads();
  const unsigned long long n = S.n_valid;
  if (tid < 256) {
    const unsigned long long t = P.digit_total[(size_t)c * 256 + tid];
    if (n == 0 || t == n) s_skip = 1;          // (benign race: every writer writes 1)
    P.digit_total[(size_t)c * 256 + tid] = 0;  // ready for the next pass
  }
  __syncthreads();
  const int skip = s_skip;
  if (tid == 0) {
    S.src[P.pass] = S.cur;
    S.skip[P.pass] = skip;
    if (!skip) S.cur ^= 1;
  }
  if (skip) return;
  uint32_t* a = P.tile_hist + (size_t)c * 256 * P.n_tiles;
  const int64_t total = (int64_t)256 * P.n_tiles;
  for (int64_t base = 0; base < total; base += 1024 * 4) {
    uint32_t v[4], run = 0;
    const int64_t i0 = base + (int64_t)tid * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = (i0 + k < total) ? a[i0 + k] : 0u; run += v[k]; }
    uint32_t inc = run;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(ANV_FULL, inc, o);
      if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
      uint32_t w = wsum[tid], wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(ANV_FULL, wi, o);
        if (tid >= o) wi += t;
      }
      wsum[tid] = wi - w;  // exclusive
    }
    __syncthreads();
    uint32_t ex = s_carry + wsum[tid >> 5] + inc - run;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (i0 + k < total) a[i0 + k] = ex; ex += v[k]; }
    __syncthreads();
    if (tid == 1023) s_carry = ex;
    __syncthreads();
  }
}

// ---- pass step 3: stable scatter ---------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK) sort_scatter_kernel(const SortParams<K> P) {
  const int c = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ColState& S = P.state[c];
  if (S.skip[P.pass]) return;
  const int64_t n = (int64_t)S.n_valid;
  const int64_t t0 = (int64_t)tile * SORT_TILE;
  if (t0 >= n) return;
  const int src = S.src[P.pass];
  const K* __restrict__ in = P.buf[src] + (size_t)c * P.stride;
  K* __restrict__ out = P.buf[src ^ 1] + (size_t)c * P.stride;
  __shared__ uint32_t wcnt[ANV_WARPS][256];
  __shared__ uint32_t gbase[256];
  for (int i = tid; i < ANV_WARPS * 256; i += ANV_BLOCK) (&wcnt[0][0])[i] = 0;
  gbase[tid] = P.tile_hist[((size_t)c * 256 + tid) * P.n_tiles + tile];
  __syncthreads();
  // warp w owns the contiguous segment [t0 + w*512, +512): rounds of 32 consecutive keys
  constexpr int WR = SORT_TILE / ANV_WARPS / 32;  // 16 rounds per warp
  K key[WR];
  uint32_t pos[WR];
  const int64_t w0 = t0 + (int64_t)warp * (SORT_TILE / ANV_WARPS);
#pragma unroll
  for (int r = 0; r < WR; ++r) {
    const int64_t i = w0 + r * 32 + lane;
    const bool ok = i < n;
    const uint32_t act = __ballot_sync(ANV_FULL, ok);
    pos[r] = 0;
    key[r] = 0;
    if (ok) {
      key[r] = in[i];
      const uint32_t d = digit_of(key[r], P.pass);
      const uint32_t m = __match_any_sync(act, d);
      const uint32_t before = wcnt[warp][d];
      pos[r] = before + __popc(m & ((1u << lane) - 1u));
      __syncwarp(m);
      if (lane == __ffs(m) - 1) wcnt[warp][d] = before + __popc(m);
    }
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive prefix over warps for digit `tid`
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < ANV_WARPS; ++w) { const uint32_t t = wcnt[w][tid]; wcnt[w][tid] = acc; acc += t; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < WR; ++r) {
    const int64_t i = w0 + r * 32 + lane;
    if (i < n) {
      const uint32_t d = digit_of(key[r], P.pass);
      out[(size_t)gbase[d] + wcnt[warp][d] + pos[r]] = key[r];
    }
  }
}

// ---- run summary of the sorted keys --------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(ANV_BLOCK) run_tile_kernel(const SortParams<K> P) {
  const int c = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const ColState& S = P.state[c];
  const int64_t n = (int64_t)S.n_valid;
  const int64_t t0 = (int64_t)tile * SORT_TILE;
  TileSummary<K>& out = P.summ[(size_t)c * P.n_tiles + tile];
  if (t0 >= n) { if (tid == 0) out.n = 0; return; }
  const int nt = (int)min((int64_t)SORT_TILE, n - t0);
  const K* __restrict__ keys = P.buf[S.cur] + (size_t)c * P.stride + t0;
  __shared__ K sk[SORT_TILE];
  __shared__ unsigned long long s_best;
  __shared__ uint32_t s_heads, s_prefix, s_suffix;
  for (int i = tid; i < nt; i += ANV_BLOCK) sk[i] = keys[i];
  if (tid == 0) { s_best = 0; s_heads = 0; s_prefix = 0; s_suffix = 0; }
  __syncthreads();
  uint32_t heads = 0;
  unsigned long long best = 0;
  for (int i = tid; i < nt; i += ANV_BLOCK) {
    if (i > 0 && sk[i] == sk[i - 1]) continue;  // not a run head
    heads += i > 0;
    int e = i + 1;                              // end (exclusive) of the run starting at i
    if (e < nt && sk[e] == sk[i]) {             // upper bound by binary search (keys are sorted)
      int lo = e, hi = nt;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sk[mid] == sk[i]) lo = mid + 1; else hi = mid; }
      e = lo;
    }
    const uint32_t len = (uint32_t)(e - i);
    if (i == 0) s_prefix = len;
    if (e == nt) s_suffix = len;
    if (i > 0 && e < nt) {
      const unsigned long long v = ((unsigned long long)len << 32) | (uint32_t)(SORT_TILE - i);  // max len, then smallest i
      best = v > best ? v : best;
    }
  }
  if (heads) atomicAdd(&s_heads, heads);
  if (best) atomicMax(&s_best, best);
  __syncthreads();
  if (tid == 0) {
    out.n = nt;
    out.first_key = sk[0];
    out.last_key = sk[nt - 1];
    out.prefix_len = s_prefix;
    out.suffix_len = s_suffix;
    out.heads_inside = s_heads;
    out.best_len = (uint32_t)(s_best >> 32);
    out.best_key = s_best ? sk[SORT_TILE - (int)(uint32_t)s_best] : (K)0;
  }
}

template <typename K>
__global__ void run_merge_kernel(const SortParams<K> P, double* mode_value, int64_t* mode_rows, int64_t* n_distinct) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_cols) return;
  const ColState& S = P.state[c];
  const int64_t n = (int64_t)S.n_valid;
  const int dt = P.cols[c].dtype;
  if (n == 0) { mode_value[c] = nan(""); mode_rows[c] = 0; n_distinct[c] = 0; return; }
  const TileSummary<K>* T = P.summ + (size_t)c * P.n_tiles;
  const int tiles = (int)((n + SORT_TILE - 1) / SORT_TILE);
  K best_key = 0, carry_key = 0, prev_last = 0;
  int64_t best_len = 0, carry_len = 0, distinct = 0;
  auto close = [&](K k, int64_t len) {  // runs are closed in ascending key order: strict > keeps the smallest key
    if (len > best_len) { best_len = len; best_key = k; }
  };
  for (int t = 0; t < tiles; ++t) {
    const TileSummary<K> s = T[t];
    distinct += s.heads_inside + ((t == 0 || s.first_key != prev_last) ? 1 : 0);
    const bool single = s.prefix_len == s.n;
    if (carry_len && carry_key == s.first_key) {
      carry_len += s.prefix_len;
    } else {
      if (carry_len) close(carry_key, carry_len);
      carry_key = s.first_key;
      carry_len = s.prefix_len;
    }
    if (!single) {
      close(carry_key, carry_len);
      if (s.best_len) close(s.best_key, s.best_len);
      carry_key = s.last_key;
      carry_len = s.suffix_len;
    }
    prev_last = s.last_key;
  }
  if (carry_len) close(carry_key, carry_len);
  const uint64_t k64 = sizeof(K) == 8 ? (uint64_t)best_key : ((uint64_t)best_key << 32);
  mode_value[c] = sorted_key_to_double(k64, dt);
  mode_rows[c] = best_len;
  n_distinct[c] = distinct;
}

template <typename K> struct Layout {
  size_t state, buf0, buf1, tile_hist, digit_total, summ, total;
  Layout(int n_cols, int64_t n_rows) {
    const int64_t stride = (n_rows + 63) & ~(int64_t)63;
    const int64_t n_tiles = (n_rows + SORT_TILE - 1) / SORT_TILE;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
    state = take((size_t)n_cols * sizeof(ColState));
    buf0 = take((size_t)n_cols * stride * sizeof(K));
    buf1 = take((size_t)n_cols * stride * sizeof(K));
    tile_hist = take((size_t)n_cols * 256 * (n_tiles > 0 ? n_tiles : 1) * 4);
    digit_total = take((size_t)n_cols * 256 * 8);
    summ = take((size_t)n_cols * (n_tiles > 0 ? n_tiles : 1) * sizeof(TileSummary<K>));
    total = o + 256;
  }
};

template <typename K>
static int run_mode_distinct(const anv_column_t* cols, int n_cols, int64_t n_rows, double* mode_value, int64_t* mode_rows,
                             int64_t* n_distinct, void* workspace, size_t workspace_bytes, cudaStream_t st) {
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
  P.digit_total = reinterpret_cast<unsigned long long*>(w + L.digit_total);
  P.summ = reinterpret_cast<TileSummary<K>*>(w + L.summ);
  ANV_CUDA(cudaMemsetAsync(P.state, 0, (size_t)n_cols * sizeof(ColState), st));
  ANV_CUDA(cudaMemsetAsync(P.digit_total, 0, (size_t)n_cols * 256 * 8, st));
  if (n_rows > 0) {
    int pack_blocks = (int)((n_rows + ANV_BLOCK * 8 - 1) / (ANV_BLOCK * 8));
    if (pack_blocks > 148 * 8) pack_blocks = 148 * 8;
    pack_kernel<K><<<dim3(pack_blocks, n_cols), ANV_BLOCK, 0, st>>>(P);
    ANV_CUDA(cudaGetLastError());
    dim3 grid(P.n_tiles, n_cols);
    for (int pass = 0; pass < (int)sizeof(K); ++pass) {
      P.pass = pass;
      sort_hist_kernel<K><<<grid, ANV_BLOCK, 0, st>>>(P);
      sort_scan_kernel<K><<<n_cols, 1024, 0, st>>>(P);
      sort_scatter_kernel<K><<<grid, ANV_BLOCK, 0, st>>>(P);
      ANV_CUDA(cudaGetLastError());
    }
    run_tile_kernel<K><<<grid, ANV_BLOCK, 0, st>>>(P);
    ANV_CUDA(cudaGetLastError());
  }
  run_merge_kernel<K><<<(n_cols + 31) / 32, 32, 0, st>>>(P, mode_value, mode_rows, n_distinct);
  ANV_CUDA(cudaGetLastError());
  return ANV_OK;
}

}  // namespace anv

using namespace anv;

extern "C" size_t anv_mode_distinct_workspace_bytes(int n_cols, int64_t n_rows, int key_bits) {
  if (n_cols <= 0 || n_rows < 0) return 256;
  return key_bits == 32 ? Layout<uint32_t>(n_cols, n_rows).total : Layout<uint64_t>(n_cols, n_rows).total;
}

extern "C" int anv_mode_distinct(const anv_column_t* cols, int n_cols, int64_t n_rows, int key_bits, double* mode_value,
                                 int64_t* mode_rows, int64_t* n_distinct, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  if (n_cols < 0 || n_rows < 0 || (key_bits != 32 && key_bits != 64)) { set_error("anv_mode_distinct: bad arguments"); return ANV_ERR_INVALID; }
  if (n_cols == 0) return ANV_OK;
  if (n_cols > 65535) { set_error("n_cols > 65535"); return ANV_ERR_UNSUPPORTED; }
  if (n_rows >= ((int64_t)1 << 32)) { set_error("anv_mode_distinct: n_rows >= 2^32 per call is not supported"); return ANV_ERR_UNSUPPORTED; }
  if (!cols || !mode_value || !mode_rows || !n_distinct || !workspace) { set_error("anv_mode_distinct: NULL argument"); return ANV_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (key_bits == 32) return run_mode_distinct<uint32_t>(cols, n_cols, n_rows, mode_value, mode_rows, n_distinct, workspace, workspace_bytes, st);
  return run_mode_distinct<uint64_t>(cols, n_cols, n_rows, mode_value, mode_rows, n_distinct, workspace, workspace_bytes, st);
}
