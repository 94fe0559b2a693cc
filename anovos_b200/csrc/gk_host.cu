// Spark's Greenwald-Khanna sketch for ONE partition that is larger than its head buffer - host code.
//
// Dataset.summary() / approxQuantile (reference stats_generator.py:488,813,908; transformers.py:215;
// quality_checker.py:845,883) answer from org.apache.spark.sql.catalyst.util.QuantileSummaries (un-vendored, Spark >= 3.1,
// restated from its published algorithm).  A partition with fewer than 50 000 non-null values keeps everything in the head
// buffer until the final compress(): its samples sit at data-independent positions (anovos_b200/shared/gk.py).  A LARGER
// partition flushes the head buffer every 50 000 insertions - sort the buffer, merge it into the sample list with
// g = 1, delta = floor(2 eps currentCount) (0 for a new minimum and for the very last element) - and compresses from the
// tail whenever 10 000 or more samples are held; the final compress() inserts what is left in the buffer and compresses
// once more.  Which samples survive depends on the ARRIVAL ORDER of the values, 50 000 at a time.
//
// The device does the heavy part: the batches are sorted by the radix-sort kernels (each batch is handed to
// anv_mode_distinct as one "column" and read back through its rank outputs).  What remains is this strictly sequential
// merge / compress over the sorted batches: ~2 steps per value, a few ns each.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace {

struct Sample { double v; long long g, d; };

// QuantileSummaries.compressImmut
void compress(std::vector<Sample>& s, double merge_threshold, std::vector<Sample>& scratch) {
  if (s.empty()) return;
  scratch.clear();
  Sample head = s.back();
  for (long long i = (long long)s.size() - 2; i >= 1; --i) {
    const Sample& x = s[(size_t)i];
    if ((double)(x.g + head.g + head.d) < merge_threshold) {
      head.g += x.g;
    } else {
      scratch.push_back(head);
      head = x;
    }
  }
  scratch.push_back(head);
  if (s.front().v <= head.v && s.size() > 1) scratch.push_back(s.front());   // "if necessary, add the minimum element"
  s.assign(scratch.rbegin(), scratch.rend());
}

// QuantileSummaries.withHeadBufferInserted for one SORTED head buffer
void insert_sorted(std::vector<Sample>& s, long long& count, const double* batch, long long m, double eps, std::vector<Sample>& out) {
  out.clear();
  out.reserve(s.size() + (size_t)m);
  size_t si = 0;
  long long cur = count;
  for (long long oi = 0; oi < m; ++oi) {
    const double x = batch[oi];
    while (si < s.size() && s[si].v <= x) out.push_back(s[si++]);
    ++cur;
    const bool first_or_last = out.empty() || (si == s.size() && oi == m - 1);
    out.push_back(Sample{x, 1, first_or_last ? 0 : (long long)floor(2.0 * eps * (double)cur)});
  }
  while (si < s.size()) out.push_back(s[si++]);
  s.swap(out);
  count = cur;
}

}  // namespace

// sorted_batches: n_values doubles = the partition's non-null values in arrival order, cut into consecutive batches of
// head_size values (the last one shorter), EACH BATCH SORTED ascending (NaN last).  Writes the compressed sketch
// (value, g, delta per sample) and returns the number of samples, or a negative anv_status (-2: capacity too small).
extern "C" long long anv_gk_partition_sketch(const double* sorted_batches, long long n_values, long long head_size, double eps,
                                             long long compress_threshold, double* out_value, long long* out_g,
                                             long long* out_delta, long long capacity) {
  if (n_values < 0 || head_size < 1 || !(eps > 0.0) || compress_threshold < 1 || (n_values > 0 && !sorted_batches)) {
    anv::set_error("anv_gk_partition_sketch: bad arguments");
    return ANV_ERR_INVALID;
  }
  std::vector<Sample> s, a, b;
  long long count = 0;
  for (long long b0 = 0; b0 < n_values; b0 += head_size) {
    const long long m = (n_values - b0 < head_size) ? n_values - b0 : head_size;
    insert_sorted(s, count, sorted_batches + b0, m, eps, a);
    if (m == head_size && (long long)s.size() >= compress_threshold) compress(s, 2.0 * eps * (double)count, b);
  }
  compress(s, 2.0 * eps * (double)count, b);          // the final compress()
  if ((long long)s.size() > capacity) { anv::set_error("anv_gk_partition_sketch: %zu samples, capacity %lld", s.size(), capacity); return ANV_ERR_WORKSPACE; }
  for (size_t i = 0; i < s.size(); ++i) { out_value[i] = s[i].v; out_g[i] = s[i].g; out_delta[i] = s[i].d; }
  return (long long)s.size();
}
