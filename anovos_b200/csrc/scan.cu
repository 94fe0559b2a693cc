// K1 / K2: the fused streaming column scan.
//
// One CTA handles one (column, row-tile): it streams the tile with 128-bit
// no-allocate loads (coalesced: consecutive threads read consecutive 16 B), keeps
//   K1  count / nonzero / min / max in native-type lanes and the pivot-shifted power
//       sums  sum d, d^2, d^3, d^4  (d = double(x) - pivot) in FP64 registers,
//   K2  the bin id from ONE fused multiply-add guess fixed up by ONE exact native-type
//       threshold compare (or a branch-free binary search), counted in per-thread
//       PRIVATE shared-memory counters (bank = lane: conflict-free, no atomics),
// then reduces with warp shuffles + one shared-memory stage and emits a mergeable
// partial per tile.  A second tiny kernel Pebay-merges the tile partials of each
// column in a fixed order, so results are run-to-run bit-stable.
//
// Replaces (reference, /root/reference/src/main/anovos): the Spark summary()/agg
// scans of data_analyzer/stats_generator.py:163,240-241,310,488,813,908,993, the
// min/max agg of data_transformer/transformers.py:217-219, the Python UDF
// bucket_label transformers.py:248-280 and the groupBy counts of
// drift_stability/drift_detector.py:252-264.
#include "common.cuh"

namespace anv {

struct Partial {  // one per (column, tile); 64 B
  int64_t n, nz;
  double mn, mx;
  double mean, m2, m3, m4;
};

struct ScanParams {
  const anv_column_t* cols;
  int n_cols;
  int64_t n_rows;
  int tile_rows;  // multiple of 1024
  // K1
  Partial* partials;
  int tiles_per_col;
  // K2
  const anv_binspec_t* specs;
  const uint64_t* cuts;
  const int32_t* card;  // codes mode: cardinality per column
  unsigned long long* counts;
  int count_stride;
  int thr_slots;  // P + 1 threshold slots reserved in shared memory
  // bin-id materialisation
  int32_t* out_bins;
  int64_t out_stride;
};

constexpr int UNROLL = 4;  // 128-bit loads in flight per thread

// ---- bin lookup ---------------------------------------------------------------------
// S[0] = lowest, S[1..B-1] = thresholds theta_0..theta_{B-2}, S[B..P] = highest.
template <typename T> struct Binner {
  const T* S;
  int B, mode, P;
  T lo, invw;

  __device__ __forceinline__ int bin0(T x) const {  // 0-based bin == #(theta_i < x)
    int k;
    if (mode == 1) {
      k = guess(x);
    } else if (mode == 2) {
      return (int)x;  // dictionary code
    } else {
      k = 0;
      const T* G = S + 1;
#pragma unroll 1
      for (int s = P >> 1; s > 0; s >>= 1) k += (G[k + s - 1] < x) ? s : 0;
    }
    if (Traits<T>::is_nan(x)) k = B - 1;  // `NaN <= c` is False for every cutoff (transformers.py:252-255)
    return k;
  }
  __device__ __forceinline__ int guess(T x) const;
};
template <> __device__ __forceinline__ int Binner<float>::guess(float x) const {
  float t = (x - lo) * invw;
  t = fminf(fmaxf(t, 0.0f), (float)(B - 1));
  const int r = __float_as_int(t + 12582912.0f) - 0x4B400000;  // round-to-nearest int, ALU only
  return max(r - 1 + (x > S[r] ? 1 : 0), 0);
}
template <> __device__ __forceinline__ int Binner<double>::guess(double x) const {
  double t = (x - lo) * invw;
  t = fmin(fmax(t, 0.0), (double)(B - 1));
  const int r = __double2loint(t + 6755399441055744.0);
  return max(r - 1 + (x > S[r] ? 1 : 0), 0);
}
template <> __device__ __forceinline__ int Binner<int32_t>::guess(int32_t) const { return 0; }
template <> __device__ __forceinline__ int Binner<int64_t>::guess(int64_t) const { return 0; }

template <typename T> __device__ __forceinline__ T cut_as(uint64_t raw);
template <> __device__ __forceinline__ float cut_as<float>(uint64_t raw) { return __uint_as_float((uint32_t)raw); }
template <> __device__ __forceinline__ int32_t cut_as<int32_t>(uint64_t raw) { return (int32_t)(uint32_t)raw; }
template <> __device__ __forceinline__ double cut_as<double>(uint64_t raw) { return __longlong_as_double((long long)raw); }
template <> __device__ __forceinline__ int64_t cut_as<int64_t>(uint64_t raw) { return (int64_t)raw; }

// ---- the tile body ------------------------------------------------------------------
template <typename T, bool MOM, int HPATH, bool ASSIGN, bool NULLS>  // HPATH: -1 none, 0 private, 1 CTA atomics, 2 global atomics
__device__ __forceinline__ void scan_tile(const ScanParams& P, const anv_column_t& col, int c, unsigned char* smem) {
  constexpr bool HIST = HPATH >= 0;
  constexpr int VEC = Traits<T>::VEC;
  constexpr uint32_t VMASK = (1u << VEC) - 1u;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * P.tile_rows;
  const int64_t r1 = min(r0 + (int64_t)P.tile_rows, P.n_rows);
  const T* __restrict__ data = reinterpret_cast<const T*>(col.data);
  const uint32_t* __restrict__ vbits = col.validity;

  // ---- K2 setup: thresholds + counters in shared memory ----------------------------
  Binner<T> bn;
  uint32_t* cnt = nullptr;
  int n_slots = 0;
  if (HIST || ASSIGN) {
    T* S = reinterpret_cast<T*>(smem);
    cnt = reinterpret_cast<uint32_t*>(smem + (size_t)P.thr_slots * 8);
    if (P.card) {  // dictionary codes: slot = code + 1
      bn.mode = 2; bn.B = P.card[c]; bn.P = 2; bn.S = S; bn.lo = 0; bn.invw = 0;
    } else {
      const anv_binspec_t sp = P.specs[c];
      bn.B = sp.n_bins; bn.mode = sp.mode; bn.S = S;
      int p2 = 2;
      while (p2 < bn.B) p2 <<= 1;
      bn.P = p2;
      bn.lo = (T)sp.lo; bn.invw = (T)sp.inv_w;
      for (int i = tid; i <= p2; i += ANV_BLOCK) {
        T v;
        if (i == 0) v = Traits<T>::lowest();
        else if (i <= bn.B - 1) v = cut_as<T>(P.cuts[sp.cut_offset + i - 1]);
        else v = Traits<T>::highest();
        S[i] = v;
      }
    }
    n_slots = bn.B + 1;
    if (HIST) {
      if (HPATH == 0) {
        for (int i = tid; i < n_slots * ANV_BLOCK; i += ANV_BLOCK) cnt[i] = 0;
      } else if (HPATH == 1) {
        for (int i = tid; i < n_slots; i += ANV_BLOCK) cnt[i] = 0;
      }
    }
    __syncthreads();
  }

  // ---- K1 setup: pivot = first finite non-null value among the tile's first 32 rows --
  double pivot = 0.0;
  if (MOM) {
    const int64_t pr = r0 + lane;
    bool ok = pr < r1;
    double pv = 0.0;
    if (ok) {
      pv = Traits<T>::to_double(data[pr]);
      if (NULLS) ok = (vbits[pr >> 5] >> (pr & 31)) & 1u;
      ok = ok && isfinite(pv);
    }
    const uint32_t m = __ballot_sync(ANV_FULL, ok);
    const int src = m ? __ffs(m) - 1 : 0;
    const int lo_ = __shfl_sync(ANV_FULL, __double2loint(pv), src);
    const int hi_ = __shfl_sync(ANV_FULL, __double2hiint(pv), src);
    pivot = m ? __hiloint2double(hi_, lo_) : 0.0;
  }

  double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
  T mn = Traits<T>::highest(), mx = Traits<T>::lowest();
  uint32_t n_ok = 0, n_nz = 0;

  auto elem = [&](T x, bool valid) -> int {
    int slot = 0;
    if (MOM) {
      double d = Traits<T>::to_double(x) - pivot;
      if (NULLS) d = valid ? d : 0.0;
      const double d2 = d * d;
      s1 += d;
      s2 += d2;
      s3 = fma(d2, d, s3);
      s4 = fma(d2, d2, s4);
      if (NULLS) {
        mn = valid ? min(mn, x) : mn;
        mx = valid ? max(mx, x) : mx;
        n_nz += (valid && x != (T)0) ? 1u : 0u;
      } else {
        mn = min(mn, x);
        mx = max(mx, x);
        n_nz += (x != (T)0) ? 1u : 0u;
      }
    }
    if (HIST || ASSIGN) {
      slot = bn.bin0(x) + 1;
      if (bn.mode == 2) slot = min(max(slot, 1), n_slots - 1);  // out-of-range code: clamp, never scribble
      if (NULLS) slot = valid ? slot : 0;
      if (HIST) {
        if (HPATH == 0) cnt[slot * ANV_BLOCK + tid] += 1;
        else if (HPATH == 1) atomicAdd(&cnt[slot], 1u);
        else atomicAdd(&P.counts[(size_t)c * P.count_stride + slot], 1ull);
      }
    }
    return slot;
  };
  int32_t* __restrict__ obins = ASSIGN ? P.out_bins + (size_t)c * P.out_stride : nullptr;
  auto store_bins = [&](int64_t row, const int (&sl)[VEC]) {
    if (VEC == 4) *reinterpret_cast<int4*>(obins + row) = make_int4(sl[0], sl[1], sl[2], sl[3]);
    else *reinterpret_cast<int2*>(obins + row) = make_int2(sl[0], sl[VEC - 1]);
  };

  // ---- stream the tile ------------------------------------------------------------
  const int64_t nvec = (r1 - r0) / VEC;  // full 16-byte vectors in this tile
  const uint4* __restrict__ vdata = reinterpret_cast<const uint4*>(data + r0);
  int64_t base = 0;
  for (; base + (int64_t)ANV_BLOCK * UNROLL <= nvec; base += (int64_t)ANV_BLOCK * UNROLL) {
    uint4 q[UNROLL];
    uint32_t vb[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t j = base + u * ANV_BLOCK + tid;
      q[u] = ldg_stream(vdata + j);
      if (NULLS) {
        const int64_t row = r0 + j * VEC;
        vb[u] = (__ldg(vbits + (row >> 5)) >> (row & 31)) & VMASK;
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      T e[VEC];
      unpack<T>(q[u], e);
      const int64_t row = r0 + (base + u * ANV_BLOCK + tid) * VEC;
      if (NULLS && MOM) n_ok += __popc(vb[u]);
      int sl[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) sl[i] = elem(e[i], NULLS ? ((vb[u] >> i) & 1u) : true);
      if (ASSIGN) store_bins(row, sl);
    }
  }
  for (int64_t j = base + tid; j < nvec; j += ANV_BLOCK) {  // remainder vectors
    const uint4 q = ldg_stream(vdata + j);
    const int64_t row = r0 + j * VEC;
    uint32_t vb = VMASK;
    if (NULLS) vb = (__ldg(vbits + (row >> 5)) >> (row & 31)) & VMASK;
    T e[VEC];
    unpack<T>(q, e);
    if (NULLS && MOM) n_ok += __popc(vb);
    int sl[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) sl[i] = elem(e[i], (vb >> i) & 1u);
    if (ASSIGN) store_bins(row, sl);
  }
  if (tid == 0) {  // scalar tail (< VEC rows, last tile only)
    for (int64_t row = r0 + nvec * VEC; row < r1; ++row) {
      bool valid = true;
      if (NULLS) valid = (vbits[row >> 5] >> (row & 31)) & 1u;
      if (NULLS && MOM) n_ok += valid;
      const int sl = elem(data[row], valid);
      if (ASSIGN) obins[row] = sl;
    }
  }

  // ---- K2 tile epilogue: fold the private counters, one integer atomic per slot -------
  if (HIST && HPATH != 2) {
    __syncthreads();
    unsigned long long* out = P.counts + (size_t)c * P.count_stride;
    if (HPATH == 0) {
      for (int s = warp; s < n_slots; s += ANV_WARPS) {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < ANV_BLOCK / 32; ++i) v += cnt[s * ANV_BLOCK + i * 32 + lane];
        v = __reduce_add_sync(ANV_FULL, v);
        if (lane == 0 && v) atomicAdd(out + s, (unsigned long long)v);
      }
    } else {
      for (int s = tid; s < n_slots; s += ANV_BLOCK) {
        const uint32_t v = cnt[s];
        if (v) atomicAdd(out + s, (unsigned long long)v);
      }
    }
  }

  // ---- K1 tile epilogue: block reduce, convert to central form, write the partial ------
  if (MOM) {
    if (!NULLS) {  // every row in range is valid: count analytically
      const int64_t mine_full = (nvec > tid) ? (nvec - tid + ANV_BLOCK - 1) / ANV_BLOCK : 0;
      n_ok = (uint32_t)(mine_full * VEC) + (tid == 0 ? (uint32_t)((r1 - r0) - nvec * VEC) : 0u);
    }
    double dmn = Traits<T>::to_double(mn), dmx = Traits<T>::to_double(mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += shfl_down_d(s1, o);
      s2 += shfl_down_d(s2, o);
      s3 += shfl_down_d(s3, o);
      s4 += shfl_down_d(s4, o);
      dmn = fmin(dmn, shfl_down_d(dmn, o));
      dmx = fmax(dmx, shfl_down_d(dmx, o));
      n_ok += __shfl_down_sync(ANV_FULL, n_ok, o);
      n_nz += __shfl_down_sync(ANV_FULL, n_nz, o);
    }
    __shared__ double red[ANV_WARPS][6];
    __shared__ uint32_t redn[ANV_WARPS][2];
    if (lane == 0) {
      red[warp][0] = s1; red[warp][1] = s2; red[warp][2] = s3; red[warp][3] = s4;
      red[warp][4] = dmn; red[warp][5] = dmx;
      redn[warp][0] = n_ok; redn[warp][1] = n_nz;
    }
    __syncthreads();
    if (tid == 0) {
      double t1 = 0, t2 = 0, t3 = 0, t4 = 0, a = INFINITY, b = -INFINITY;
      int64_t n = 0, nz = 0;
#pragma unroll
      for (int w = 0; w < ANV_WARPS; ++w) {  // fixed order: deterministic
        t1 += red[w][0]; t2 += red[w][1]; t3 += red[w][2]; t4 += red[w][3];
        a = fmin(a, red[w][4]); b = fmax(b, red[w][5]);
        n += redn[w][0]; nz += redn[w][1];
      }
      Partial out;
      out.n = n; out.nz = nz; out.mn = a; out.mx = b;
      if (n > 0) {
        const double dn = (double)n;
        const double dl = t1 / dn;  // mean - pivot
        out.mean = pivot + dl;
        out.m2 = t2 - t1 * dl;
        out.m3 = t3 - 3.0 * dl * t2 + 2.0 * dl * dl * t1;
        out.m4 = t4 - 4.0 * dl * t3 + 6.0 * dl * dl * t2 - 3.0 * dl * dl * dl * t1;
      } else {
        out.mean = 0.0; out.m2 = out.m3 = out.m4 = 0.0;
      }
      P.partials[(size_t)c * P.tiles_per_col + blockIdx.x] = out;
    }
  }
}

template <bool MOM, int HPATH, bool ASSIGN>
__global__ void __launch_bounds__(ANV_BLOCK) scan_kernel(const ScanParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int c = blockIdx.y;
  const anv_column_t col = P.cols[c];
#define ANV_DISPATCH(T)                                                          \
  if (col.validity) scan_tile<T, MOM, HPATH, ASSIGN, true>(P, col, c, smem);     \
  else scan_tile<T, MOM, HPATH, ASSIGN, false>(P, col, c, smem);
  switch (col.dtype) {
    case ANV_F32: ANV_DISPATCH(float) break;
    case ANV_F64: ANV_DISPATCH(double) break;
    case ANV_I32: ANV_DISPATCH(int32_t) break;
    case ANV_I64: ANV_DISPATCH(int64_t) break;
    default: break;
  }
#undef ANV_DISPATCH
}

// One warp per column: lane-strided sequential Pebay merge, then a shuffle tree.
__global__ void __launch_bounds__(32) finalize_moments(const Partial* partials, int tiles_per_col, anv_moments_t* out) {
  const int c = blockIdx.x, lane = threadIdx.x;
  const Partial* p = partials + (size_t)c * tiles_per_col;
  Central acc{0, 0, 0, 0, 0};
  int64_t n = 0, nz = 0;
  double mn = INFINITY, mx = -INFINITY;
  for (int t = lane; t < tiles_per_col; t += 32) {
    const Partial q = p[t];
    if (q.n > 0) {
      acc = merge_central(acc, Central{(double)q.n, q.mean, q.m2, q.m3, q.m4});
      mn = fmin(mn, q.mn);
      mx = fmax(mx, q.mx);
    }
    n += q.n;
    nz += q.nz;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Central b;
    b.n = shfl_down_d(acc.n, o); b.mean = shfl_down_d(acc.mean, o);
    b.m2 = shfl_down_d(acc.m2, o); b.m3 = shfl_down_d(acc.m3, o); b.m4 = shfl_down_d(acc.m4, o);
    acc = merge_central(acc, b);
    mn = fmin(mn, shfl_down_d(mn, o));
    mx = fmax(mx, shfl_down_d(mx, o));
    n += shfl_down_ll(n, o);
    nz += shfl_down_ll(nz, o);
  }
  if (lane == 0) {
    anv_moments_t r;
    r.n_valid = n; r.n_nonzero = nz;
    if (n > 0) { r.min = mn; r.max = mx; r.mean = acc.mean; r.m2 = acc.m2; r.m3 = acc.m3; r.m4 = acc.m4; }
    else { r.min = r.max = r.mean = nan(""); r.m2 = r.m3 = r.m4 = 0.0; }
    out[c] = r;
  }
}

// ---- host side ----------------------------------------------------------------------
static int pick_tile_rows(int64_t n_rows, int n_cols) {
  // >= ~8 tiles per SM across the launch, tile in [16Ki, 256Ki] rows, multiple of 1024
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t want_tiles = (int64_t)sms * 8;
  int64_t per_col = (want_tiles + n_cols - 1) / (n_cols > 0 ? n_cols : 1);
  int64_t tr = per_col > 0 ? n_rows / per_col : n_rows;
  int64_t t = 16384;
  while (t < tr && t < 262144) t <<= 1;
  return (int)t;
}

static size_t hist_smem(int count_stride, int* path, int* thr_slots, bool codes = false) {
  int nb = count_stride - 1, p2 = 2;
  while (p2 < nb) p2 <<= 1;
  *thr_slots = codes ? 2 : p2 + 2;  // dictionary codes need no thresholds
  size_t thr = (size_t)(*thr_slots) * 8;
  if (count_stride <= 40) { *path = 0; return thr + (size_t)count_stride * ANV_BLOCK * 4; }
  if (count_stride <= 10240) { *path = 1; return thr + (size_t)count_stride * 4; }
  *path = 2;
  return thr;
}

template <bool MOM, int HPATH, bool ASSIGN>
static int launch_scan(ScanParams& P, size_t smem, cudaStream_t st) {
  if (P.n_cols <= 0 || P.n_rows <= 0) return ANV_OK;
  dim3 grid((unsigned)((P.n_rows + P.tile_rows - 1) / P.tile_rows), (unsigned)P.n_cols);
  if (smem > 48 * 1024)
    ANV_CUDA(cudaFuncSetAttribute(scan_kernel<MOM, HPATH, ASSIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scan_kernel<MOM, HPATH, ASSIGN><<<grid, ANV_BLOCK, smem, st>>>(P);
  ANV_CUDA(cudaGetLastError());
  return ANV_OK;
}

static int check_common(const void* cols, int n_cols, int64_t n_rows) {
  if (n_cols < 0 || n_rows < 0) { set_error("negative n_cols / n_rows"); return ANV_ERR_INVALID; }
  if (n_cols > 65535) { set_error("n_cols > 65535: split the frame into column blocks"); return ANV_ERR_UNSUPPORTED; }
  if (n_cols > 0 && !cols) { set_error("cols is NULL"); return ANV_ERR_INVALID; }
  return ANV_OK;
}

}  // namespace anv

using namespace anv;

extern "C" size_t anv_moments_workspace_bytes(int n_cols, int64_t n_rows) {
  if (n_cols <= 0 || n_rows <= 0) return 64;
  const int tr = pick_tile_rows(n_rows, n_cols);
  const int64_t tiles = (n_rows + tr - 1) / tr;
  return (size_t)tiles * (size_t)n_cols * sizeof(Partial) + 64;
}

extern "C" int anv_moments(const anv_column_t* cols, int n_cols, int64_t n_rows, anv_moments_t* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (int e = check_common(cols, n_cols, n_rows)) return e;
  if (n_cols == 0) return ANV_OK;
  if (!out) { set_error("out is NULL"); return ANV_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  ScanParams P{};
  P.cols = cols; P.n_cols = n_cols; P.n_rows = n_rows;
  P.tile_rows = pick_tile_rows(n_rows > 0 ? n_rows : 1, n_cols);
  P.tiles_per_col = (int)((n_rows + P.tile_rows - 1) / P.tile_rows);
  if (n_rows == 0) P.tiles_per_col = 0;
  if (workspace_bytes < anv_moments_workspace_bytes(n_cols, n_rows) || !workspace) {
    set_error("anv_moments_t: workspace too small (%zu < %zu)", workspace_bytes, anv_moments_workspace_bytes(n_cols, n_rows));
    return ANV_ERR_WORKSPACE;
  }
  P.partials = reinterpret_cast<Partial*>(workspace);
  if (int e = launch_scan<true, -1, false>(P, 0, st)) return e;
  finalize_moments<<<n_cols, 32, 0, st>>>(P.partials, P.tiles_per_col, out);
  ANV_CUDA(cudaGetLastError());
  return ANV_OK;
}

extern "C" int anv_hist(const anv_column_t* cols, const anv_binspec_t* specs, const void* cuts, int n_cols, int64_t n_rows,
                        uint64_t* counts, int count_stride, void* stream) {
  if (int e = check_common(cols, n_cols, n_rows)) return e;
  if (n_cols == 0) return ANV_OK;
  if (!specs || !counts || count_stride < 2) { set_error("anv_hist: bad specs/counts/count_stride"); return ANV_ERR_INVALID; }
  if (count_stride > 16385) { set_error("anv_hist: more than 16384 bins per column is not supported"); return ANV_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  ANV_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_cols * count_stride * sizeof(uint64_t), st));
  ScanParams P{};
  P.cols = cols; P.n_cols = n_cols; P.n_rows = n_rows;
  P.tile_rows = pick_tile_rows(n_rows > 0 ? n_rows : 1, n_cols);
  P.specs = specs; P.cuts = reinterpret_cast<const uint64_t*>(cuts);
  P.counts = reinterpret_cast<unsigned long long*>(counts); P.count_stride = count_stride;
  int path = 0;
  size_t smem = hist_smem(count_stride, &path, &P.thr_slots);
  if (path == 0) return launch_scan<false, 0, false>(P, smem, st);
  if (path == 1) return launch_scan<false, 1, false>(P, smem, st);
  return launch_scan<false, 2, false>(P, smem, st);
}

extern "C" int anv_hist_codes(const anv_column_t* cols, const int32_t* cardinality, int n_cols, int64_t n_rows,
                              uint64_t* counts, int count_stride, void* stream) {
  if (int e = check_common(cols, n_cols, n_rows)) return e;
  if (n_cols == 0) return ANV_OK;
  if (!cardinality || !counts || count_stride < 2) { set_error("anv_hist_codes: bad arguments"); return ANV_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  ANV_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_cols * count_stride * sizeof(uint64_t), st));
  ScanParams P{};
  P.cols = cols; P.n_cols = n_cols; P.n_rows = n_rows;
  P.tile_rows = pick_tile_rows(n_rows > 0 ? n_rows : 1, n_cols);
  P.card = cardinality;
  P.counts = reinterpret_cast<unsigned long long*>(counts); P.count_stride = count_stride;
  int path = 0;
  size_t smem = hist_smem(count_stride, &path, &P.thr_slots, true);
  if (path == 0) return launch_scan<false, 0, false>(P, smem, st);
  if (path == 1) return launch_scan<false, 1, false>(P, smem, st);
  return launch_scan<false, 2, false>(P, smem, st);
}

extern "C" int anv_moments_hist(const anv_column_t* cols, const anv_binspec_t* specs, const void* cuts, int n_cols,
                                int64_t n_rows, anv_moments_t* out, uint64_t* counts, int count_stride, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (int e = check_common(cols, n_cols, n_rows)) return e;
  if (n_cols == 0) return ANV_OK;
  if (!specs || !counts || !out || count_stride < 2) { set_error("anv_moments_hist: bad arguments"); return ANV_ERR_INVALID; }
  if (count_stride > 16385) { set_error("anv_moments_hist: more than 16384 bins per column is not supported"); return ANV_ERR_UNSUPPORTED; }
  if (workspace_bytes < anv_moments_workspace_bytes(n_cols, n_rows) || !workspace) {
    set_error("anv_moments_hist: workspace too small");
    return ANV_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ANV_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_cols * count_stride * sizeof(uint64_t), st));
  ScanParams P{};
  P.cols = cols; P.n_cols = n_cols; P.n_rows = n_rows;
  P.tile_rows = pick_tile_rows(n_rows > 0 ? n_rows : 1, n_cols);
  P.tiles_per_col = n_rows ? (int)((n_rows + P.tile_rows - 1) / P.tile_rows) : 0;
  P.partials = reinterpret_cast<Partial*>(workspace);
  P.specs = specs; P.cuts = reinterpret_cast<const uint64_t*>(cuts);
  P.counts = reinterpret_cast<unsigned long long*>(counts); P.count_stride = count_stride;
  int path = 0;
  size_t smem = hist_smem(count_stride, &path, &P.thr_slots);
  if (path == 0) {
    if (int e = launch_scan<true, 0, false>(P, smem, st)) return e;
  } else {  // wide histograms: two kernels (the fused variant only pays off with private counters)
    if (int e = launch_scan<true, -1, false>(P, 0, st)) return e;
    if (int e = (path == 1 ? launch_scan<false, 1, false>(P, smem, st) : launch_scan<false, 2, false>(P, smem, st))) return e;
  }
  finalize_moments<<<n_cols, 32, 0, st>>>(P.partials, P.tiles_per_col, out);
  ANV_CUDA(cudaGetLastError());
  return ANV_OK;
}

extern "C" int anv_bin_assign(const anv_column_t* cols, const anv_binspec_t* specs, const void* cuts, int n_cols,
                              int64_t n_rows, int max_bins, int32_t* out_bins, int64_t out_stride, void* stream) {
  if (int e = check_common(cols, n_cols, n_rows)) return e;
  if (n_cols == 0) return ANV_OK;
  if (!specs || !out_bins || out_stride < n_rows || (out_stride & 3) || max_bins < 2 || max_bins > 4096) {
    set_error("anv_bin_assign: bad arguments (out_stride must be >= n_rows and a multiple of 4; 2 <= max_bins <= 4096)");
    return ANV_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ScanParams P{};
  P.cols = cols; P.n_cols = n_cols; P.n_rows = n_rows;
  P.tile_rows = pick_tile_rows(n_rows > 0 ? n_rows : 1, n_cols);
  P.specs = specs; P.cuts = reinterpret_cast<const uint64_t*>(cuts);
  P.out_bins = out_bins; P.out_stride = out_stride;
  int path = 0;
  hist_smem(max_bins + 1, &path, &P.thr_slots);
  return launch_scan<false, -1, true>(P, (size_t)P.thr_slots * 8, st);
}
