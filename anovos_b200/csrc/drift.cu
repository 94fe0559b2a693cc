// K3: drift reduce.  One thread per column walks the (p, q) table in key order
// exactly like the reference's orderBy(i) + sum / window aggregations
// (/root/reference/src/main/anovos/drift_stability/drift_detector.py:266-335):
// fillna(1e-4) for a key missing on one side, replace(0 -> 1e-4), then
// PSI = sum (p-q) ln(p/q), HD = sqrt(sum (sqrt p - sqrt q)^2 / 2),
// JSD = (sum p ln(p/m) + sum q ln(q/m)) / 2 with m = (p+q)/2, KS = max |cum p - cum q|.
// Everything in FP64.  Narrow tables (binned numeric columns, small dictionaries): one thread per
// column, strictly sequential.  Wide tables (string columns with > DRIFT_WIDE keys): one CTA per column,
// each thread reduces a contiguous key segment, the segments are combined IN ORDER by thread 0 and the
// KS running sums restart from the exact segment offsets - deterministic (fixed segmentation), same
// sums up to FP64 re-association.
#include "common.cuh"

namespace anv {

struct DriftAcc {
  double psi = 0, hd = 0, pm = 0, qm = 0, cp = 0, cq = 0, ks = 0;
  int rows = 0;
  __device__ __forceinline__ void row(double p, double q) {
    if (p == 0.0) p = 0.0001;
    if (q == 0.0) q = 0.0001;
    psi += (p - q) * log(p / q);
    const double t = sqrt(p) - sqrt(q);
    hd += t * t;
    const double m = (p + q) / 2;
    pm += p * log(p / m);
    qm += q * log(q / m);
    cp += p;
    cq += q;
    ks = fmax(ks, fabs(cp - cq));
    ++rows;
  }
};

constexpr int DRIFT_WIDE = 96, DRIFT_THREADS = 128;

__device__ __forceinline__ bool drift_pq(const unsigned long long* s, const unsigned long long* t, const double* sp, int src_is_p,
                                         int k, double n_src, double n_tgt, double& p, double& q) {
  bool ps;
  const bool pt = t[k] > 0;
  q = pt ? (double)t[k] / n_tgt : 0.0001;
  if (src_is_p) { ps = !isnan(sp[k]); p = ps ? sp[k] : 0.0001; }
  else { ps = s[k] > 0; p = ps ? (double)s[k] / n_src : 0.0001; }
  return ps || pt;
}

__global__ void __launch_bounds__(DRIFT_THREADS) drift_reduce_wide_kernel(
    const unsigned long long* __restrict__ src, const unsigned long long* __restrict__ tgt, const double* __restrict__ src_p,
    int src_is_p, const int32_t* __restrict__ n_slots, const int32_t* __restrict__ kind, int stride, double n_src,
    double n_tgt, anv_drift_t* __restrict__ out) {
  const int c = blockIdx.x, tid = threadIdx.x;
  const int ns = n_slots[c];
  if (ns <= DRIFT_WIDE) return;
  const unsigned long long* t = tgt + (size_t)c * stride;
  const unsigned long long* s = src_is_p ? nullptr : src + (size_t)c * stride;
  const double* sp = src_is_p ? src_p + (size_t)c * stride : nullptr;
  __shared__ double part[6][DRIFT_THREADS];   // psi, hd, pm, qm, cp, cq of each segment -> cp/cq become offsets
  __shared__ int part_rows[DRIFT_THREADS];
  __shared__ double ks_part[DRIFT_THREADS];
  __shared__ DriftAcc total;
  const int per = (ns - 1 + DRIFT_THREADS - 1) / DRIFT_THREADS;
  const int k0 = 1 + tid * per, k1 = min(ns, k0 + per);
  DriftAcc a;
  for (int k = k0; k < k1; ++k) {
    double p, q;
    if (drift_pq(s, t, sp, src_is_p, k, n_src, n_tgt, p, q)) a.row(p, q);
  }
  part[0][tid] = a.psi; part[1][tid] = a.hd; part[2][tid] = a.pm; part[3][tid] = a.qm; part[4][tid] = a.cp; part[5][tid] = a.cq;
  part_rows[tid] = a.rows;
  __syncthreads();
  if (tid == 0) {
    DriftAcc g;
    const bool s_null = src_is_p ? !isnan(sp[0]) : (s[0] > 0);
    const bool t_null = t[0] > 0;
    if (kind[c] == 0) {
      if (s_null || t_null) g.row(0.0001, 0.0001);
    } else {
      if (s_null) g.row(0.0001, 0.0001);
      if (t_null) g.row(0.0001, 0.0001);
    }
    for (int i = 0; i < DRIFT_THREADS; ++i) {
      const double cp = part[4][i], cq = part[5][i];
      part[4][i] = g.cp; part[5][i] = g.cq;                 // running sums BEFORE segment i
      g.psi += part[0][i]; g.hd += part[1][i]; g.pm += part[2][i]; g.qm += part[3][i];
      g.cp += cp; g.cq += cq; g.rows += part_rows[i];
    }
    total = g;
  }
  __syncthreads();
  double cp = part[4][tid], cq = part[5][tid], ks = 0;
  for (int k = k0; k < k1; ++k) {
    double p, q;
    if (drift_pq(s, t, sp, src_is_p, k, n_src, n_tgt, p, q)) {
      if (p == 0.0) p = 0.0001;
      if (q == 0.0) q = 0.0001;
      cp += p; cq += q;
      ks = fmax(ks, fabs(cp - cq));
    }
  }
  ks_part[tid] = ks;
  __syncthreads();
  if (tid == 0) {
    double m = total.ks;                                     // the null rows (|1e-4 - 1e-4| = 0, kept for symmetry)
    for (int i = 0; i < DRIFT_THREADS; ++i) m = fmax(m, ks_part[i]);
    anv_drift_t r;
    r.n_rows = total.rows; r.reserved = 0;
    if (total.rows) { r.psi = total.psi; r.hd = sqrt(total.hd / 2); r.jsd = (total.pm + total.qm) / 2; r.ks = m; }
    else { r.psi = r.hd = r.jsd = r.ks = nan(""); }
    out[c] = r;
  }
}

__global__ void drift_reduce_kernel(const unsigned long long* __restrict__ src, const unsigned long long* __restrict__ tgt,
                                    const double* __restrict__ src_p, int src_is_p, const int32_t* __restrict__ n_slots,
                                    const int32_t* __restrict__ kind, int n_cols, int stride, double n_src, double n_tgt,
                                    anv_drift_t* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  const unsigned long long* t = tgt + (size_t)c * stride;
  const unsigned long long* s = src_is_p ? nullptr : src + (size_t)c * stride;
  const double* sp = src_is_p ? src_p + (size_t)c * stride : nullptr;
  const int ns = n_slots[c];
  if (ns > DRIFT_WIDE) return;  // handled by drift_reduce_wide_kernel
  DriftAcc a;
  // slot 0: the null group.  count(col) of a null group is 0 -> p (or q) = 0 -> 1e-4.
  const bool s_null = src_is_p ? !isnan(sp[0]) : (s[0] > 0);
  const bool t_null = t[0] > 0;
  if (kind[c] == 0) {  // binned numeric: fillna(-1) makes the null groups join on key -1 (:252-256,264)
    if (s_null || t_null) a.row(0.0001, 0.0001);
  } else {             // string keys stay SQL NULL and never match in the full outer join (:266)
    if (s_null) a.row(0.0001, 0.0001);
    if (t_null) a.row(0.0001, 0.0001);
  }
  for (int k = 1; k < ns; ++k) {
    bool ps, pt = t[k] > 0;
    double p, q = pt ? (double)t[k] / n_tgt : 0.0001;
    if (src_is_p) { ps = !isnan(sp[k]); p = ps ? sp[k] : 0.0001; }
    else { ps = s[k] > 0; p = ps ? (double)s[k] / n_src : 0.0001; }
    if (ps || pt) a.row(p, q);
  }
  anv_drift_t r;
  r.n_rows = a.rows; r.reserved = 0;
  if (a.rows) { r.psi = a.psi; r.hd = sqrt(a.hd / 2); r.jsd = (a.pm + a.qm) / 2; r.ks = a.ks; }
  else { r.psi = r.hd = r.jsd = r.ks = nan(""); }
  out[c] = r;
}

}  // namespace anv

extern "C" int anv_drift_reduce(const uint64_t* src_counts, const uint64_t* tgt_counts, const double* src_p, int src_is_p,
                                const int32_t* n_slots, const int32_t* kind, int n_cols, int count_stride, int64_t n_src,
                                int64_t n_tgt, anv_drift_t* out, void* stream) {
  if (n_cols < 0) { anv::set_error("anv_drift_reduce: negative n_cols"); return ANV_ERR_INVALID; }
  if (n_cols == 0) return ANV_OK;
  if (!tgt_counts || !n_slots || !kind || !out || (src_is_p ? !src_p : !src_counts)) {
    anv::set_error("anv_drift_reduce: NULL argument");
    return ANV_ERR_INVALID;
  }
  anv::drift_reduce_kernel<<<(n_cols + 31) / 32, 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(src_counts), reinterpret_cast<const unsigned long long*>(tgt_counts),
      src_p, src_is_p, n_slots, kind, n_cols, count_stride, (double)n_src, (double)n_tgt, out);
  ANV_CUDA(cudaGetLastError());
  if (count_stride > anv::DRIFT_WIDE) {  // some column may hold a wide table
    anv::drift_reduce_wide_kernel<<<n_cols, anv::DRIFT_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long*>(src_counts), reinterpret_cast<const unsigned long long*>(tgt_counts),
        src_p, src_is_p, n_slots, kind, count_stride, (double)n_src, (double)n_tgt, out);
    ANV_CUDA(cudaGetLastError());
  }
  return ANV_OK;
}
