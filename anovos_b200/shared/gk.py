"""Which order statistic does Spark return?  (host arithmetic only - the element itself is fetched
by the radix-select / sort kernels.)

The reference takes its percentiles from Spark's Greenwald-Khanna sketch: `Dataset.summary()` (eps 1e-4,
stats_generator.py:488,813,908) and `approxQuantile(cols, probs, 0.01)` (transformers.py:215,
quality_checker.py:845,883).  For ONE partition of fewer than 50 000 non-null values - every unit test of the
reference - the sketch never flushes its head buffer before the final compress(), so the sample positions it
keeps, and the position a query returns, are a function of (n, eps, p) only: Spark's answer is the order
statistic at a *shifted rank*.  `spark_rank` computes that rank (closed form per surviving sample, O(#samples));
with it the product reproduces e.g. all outlier_detection counts of test_quality_checker.py:526-637 and the IV / IG
pins that depend on approxQuantile cutoffs, which the textbook rank ceil(p*n) does not.

For n >= 50 000, or several partitions, Spark's answer depends on arrival order and partitioning (any element
within eps*n ranks of ceil(p*n) can come out): the exact rank max(1, ceil(p*n)) is used there, which lies inside
that band by construction.

Algorithm restated (un-vendored: org.apache.spark.sql.catalyst.util.QuantileSummaries, Spark >= 3.1):
  insert (one sorted batch): sample k (1-based) gets g = 1, delta = floor(2*eps*k), first and last delta = 0
  compress from the tail   : head absorbs its predecessor while 1 + head.g + head.delta < 2*eps*n; sample 1 is kept
  query(p)                 : p <= eps -> first, p >= 1-eps -> last, else the first sample i (not the last) with
                             maxRank_i - T <= ceil(p*n) <= minRank_i + T,  T = max(g + delta) / 2, else the last.
"""
from __future__ import annotations

import functools
import math

import numpy as np

HEAD_SIZE = 50000          # QuantileSummaries.defaultHeadSize
COMPRESS_THRESHOLD = 10000  # QuantileSummaries.defaultCompressThreshold
SUMMARY_EPS = 1e-4         # Dataset.summary(): ApproximatePercentile, accuracy 10000
APPROX_QUANTILE_EPS = 0.01  # the relativeError of every approxQuantile call on the path


@functools.lru_cache(maxsize=256)
def _summary(n: int, eps: float):
    """-> (pos [0-based positions in the sorted values], min_rank, max_rank, target_error) of the compressed sketch."""
    thr_up = int(math.ceil(2.0 * eps * n))       # 1 + g + delta < thr  <=>  1 + g + delta < ceil(thr) for integers
    pos, g = [], []
    i = n - 1                                   # current head (0-based); the last sample has delta 0
    while i >= 1:
        d = 0 if i == n - 1 else int(math.floor(2.0 * eps * (i + 1)))
        # the head (g = 1) absorbs predecessors while 1 + g + d < thr  ->  final g = max(1, ceil(thr) - 1 - d),
        # limited by the predecessors available above sample 0 (which is never absorbed)
        want = max(1, thr_up - 1 - d)
        take = min(want, i)                     # samples i, i-1, ..., i-take+1 (all >= 1)
        pos.append(i)
        g.append(take)
        i -= take
    pos.append(0)
    g.append(1)
    pos = np.array(pos[::-1], dtype=np.int64)
    g = np.array(g[::-1], dtype=np.int64)
    delta = np.floor(2.0 * eps * (pos + 1)).astype(np.int64)
    delta[0] = 0
    delta[-1] = 0
    min_rank = np.cumsum(g)
    return pos, min_rank, min_rank + delta, float((g + delta).max()) / 2.0


def spark_rank(n: int, p: float, eps) -> int:
    """1-based rank (among the n non-null values, ascending, NaN last) of the element Spark returns for quantile p.
    eps None, n >= 50 000 or n <= 0: the exact rule max(1, ceil(p*n)) with p*n in float64 (0 when n == 0)."""
    if n <= 0:
        return 0
    if eps is None or n >= HEAD_SIZE:
        return max(1, int(math.ceil(p * n)))
    if n == 1:
        return 1
    pos, min_rank, max_rank, te = _summary(int(n), float(eps))
    if p <= eps:
        return int(pos[0]) + 1
    if p >= 1 - eps:
        return int(pos[-1]) + 1
    rank = int(math.ceil(p * n))
    ok = (max_rank[:-1] - te <= rank) & (rank <= min_rank[:-1] + te)
    hit = int(np.argmax(ok)) if ok.any() else len(pos) - 1
    return int(pos[hit]) + 1


@functools.lru_cache(maxsize=4096)
def _spark_ranks(n: int, probs: tuple, eps):
    return tuple(spark_rank(n, p, eps) for p in probs)


def spark_ranks(n: int, probs, eps):
    return list(_spark_ranks(int(n), tuple(probs), eps))


# ---- several partitions: the sketches are merged in partition order (host arithmetic on a few thousand samples) ------
# A partition of fewer than 50 000 non-null values contributes the samples at the data-independent positions of
# `_summary`; their VALUES come from the sort kernel (order statistics of the partition).  Merging is value-dependent
# (QuantileSummaries.merge interleaves the two sample lists), so it happens here, on the host, on the samples only.

def partition_samples(values_at_positions, n: int, eps: float):
    """-> [(value, g, delta)] of one partition's compressed sketch; values_at_positions[i] = the order statistic at
    position sample_positions(n, eps)[i]."""
    if n <= 0:
        return []
    pos, min_rank, max_rank, _ = _summary(int(n), float(eps))
    g = np.diff(np.concatenate([[0], min_rank]))
    return list(zip([float(v) for v in values_at_positions], g.tolist(), (max_rank - min_rank).tolist()))


def sample_positions(n: int, eps: float):
    """0-based positions (in the partition's sorted non-null values) of the samples its sketch keeps."""
    if n <= 0:
        return np.zeros(0, np.int64)
    if n == 1:
        return np.zeros(1, np.int64)
    return _summary(int(n), float(eps))[0]


def _compress(samples, thr):
    if not samples:
        return []
    res = []
    hv, hg, hd = samples[-1]
    for i in range(len(samples) - 2, 0, -1):
        v, g, d = samples[i]
        if g + hg + hd < thr:
            hg += g
        else:
            res.append((hv, hg, hd))
            hv, hg, hd = v, g, d
    res.append((hv, hg, hd))
    if samples[0][0] <= hv and len(samples) > 1:
        res.append(samples[0])
    res.reverse()
    return res


def merge_samples(a, na: int, b, nb: int, eps: float):
    """QuantileSummaries.merge (Spark >= 3.0) of two compressed sketches -> (samples, count)."""
    if nb == 0:
        return list(a), na
    if na == 0:
        return list(b), nb
    add_a, add_b = int(math.floor(2 * eps * nb)), int(math.floor(2 * eps * na))
    out, i, j = [], 0, 0
    la, lb = len(a), len(b)
    while i < la and j < lb:
        if a[i][0] < b[j][0]:
            v, g, d = a[i]
            d += add_a if j > 0 else 0
            i += 1
        else:
            v, g, d = b[j]
            d += add_b if i > 0 else 0
            j += 1
        out.append((v, g, d))
    out.extend(a[i:])
    out.extend(b[j:])
    return _compress(out, 2 * eps * (na + nb)), na + nb


def query_samples(samples, n: int, eps: float, p: float):
    if not samples:
        return None
    if p <= eps:
        return samples[0][0]
    if p >= 1 - eps:
        return samples[-1][0]
    te = max(g + d for _, g, d in samples) / 2.0
    rank = int(math.ceil(p * n))
    min_rank, i = samples[0][1], 0
    last = len(samples) - 1
    while i < last:
        if min_rank + samples[i][2] - te <= rank <= min_rank + te:
            return samples[i][0]
        i += 1
        min_rank += samples[i][1]
    return samples[-1][0]
