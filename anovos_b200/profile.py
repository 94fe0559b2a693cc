"""Per-frame cache of kernel results so that the seven stats functions (and drift) scan a
frame once per kind of pass instead of once per function (the reference re-reads the input
~2C+10 times for a full stats_generator block)."""
from __future__ import annotations

import math

import numpy as np

from . import _lib, engine
from .frame import ColumnFrame
from .shared.gk import APPROX_QUANTILE_EPS, SUMMARY_EPS


SUMMARY_PROBS = [0.01, 0.05, 0.1, 0.25, 0.5, 0.75, 0.9, 0.95, 0.99]
DEFAULT_HLL_P = 9     # approx_count_distinct's default rsd = 0.05 -> p = ceil(2 log2(1.106 / 0.05)) = 9


def _cache(frame: ColumnFrame, key):
    return frame._cache.setdefault(key, {})


def moments(frame: ColumnFrame, names):
    """dict name -> numpy record (n_valid, n_nonzero, min, max, mean, m2, m3, m4)."""
    c = _cache(frame, "moments")
    todo = [n for n in names if n not in c]
    if todo:
        res = engine.moments(frame, todo)
        for n, r in zip(todo, res):
            c[n] = r
    return {n: c[n] for n in names}


def moments_table(frame: ColumnFrame, names):
    """The moment records of `names` as ONE structured array (fields as in moments()): column-wise post-processing reads a
    field of all attributes at once instead of one record at a time."""
    m = moments(frame, names)
    return np.array([m[n] for n in names], dtype=engine._MOM_DT)


def n_valid(frame: ColumnFrame, names):
    """Non-null counts.  Numeric columns: the fused moments pass.  String columns: slot 0 of their code histogram - the
    pass that mode / distinct / HLL++ / drift need anyway, so a full stats run reads a string column once, not twice."""
    mom_cache = frame._cache.get("moments", {})
    cat = [n for n in names if frame.column(n).kind == "cat" and n not in mom_cache]
    cc = code_counts(frame, cat) if cat else {}
    have = {n: frame.n_rows - int(cc[n][0]) for n in cat}
    m = moments(frame, [n for n in names if n not in have])
    return {n: have[n] if n in have else int(m[n]["n_valid"]) for n in names}


def quantiles(frame: ColumnFrame, names, probs, eps=SUMMARY_EPS):
    """dict name -> list of order statistics (None if empty) at the ranks Spark returns for `probs`: eps is the
    relativeError of the replaced call (summary(): 1e-4, approxQuantile(..., 0.01): 0.01), see shared/gk.py."""
    if getattr(frame, "spark_partitions", False) and eps is not None:
        gc = _cache(frame, ("gk_quantiles", eps))
        # resolve the nine summary() percentiles together with the request: one sort per partition serves them all
        ask = list(dict.fromkeys(list(probs) + (SUMMARY_PROBS if eps == SUMMARY_EPS else [])))
        todo = [n for n in names if any((n, p) not in gc for p in probs)]
        if todo:
            res = frame.gk_quantiles(todo, ask, eps)
            if res is not None:
                for n in todo:
                    for p, v in zip(ask, res[n]):
                        gc[(n, p)] = v
        if all((n, p) in gc for n in names for p in probs):
            return {n: [gc[(n, p)] for p in probs] for n in names}
    c = _cache(frame, "quantiles")          # name -> {1-based rank: order statistic}
    rc = _cache(frame, "quantile_ranks")
    pk = (tuple(probs), eps)
    want = {}
    mom = None
    for n in names:
        r = rc.get((n, pk))
        if r is None:
            if mom is None:
                mom = moments(frame, names)
            r = rc[(n, pk)] = engine.quantile_ranks(int(mom[n]["n_valid"]), probs, eps)
        want[n] = r
    todo = []
    for n in names:
        have = c.get(n)
        if have is None:
            have = c[n] = {}
        for r in want[n]:
            if r and r not in have:
                todo.append(n)
                break
    if todo:
        # a select pass costs the same for 1 or 16 ranks: always resolve the nine summary() percentiles too, so
        # median / IQR / percentiles share ONE radix select per column
        if mom is None:
            mom = moments(frame, names)
        sets = {}
        for n in todo:
            extra = engine.quantile_ranks(int(mom[n]["n_valid"]), SUMMARY_PROBS, SUMMARY_EPS)
            sets[n] = sorted(set(r for r in want[n] + extra if r and r not in c[n]))
        width = max(len(v) for v in sets.values())
        rk = np.zeros((len(todo), max(width, 1)), np.int64)
        for i, n in enumerate(todo):
            rs = sets[n]
            rk[i, :len(rs)] = rs
        vals = engine.select_ranks(frame, todo, rk)
        for i, n in enumerate(todo):
            k = len(sets[n])
            c[n].update(zip(sets[n], vals[i, :k].tolist()))
    out = {}
    for n in names:
        have = c[n]
        out[n] = [have[r] if r else None for r in want[n]]
    return out


def code_counts(frame: ColumnFrame, names):
    """dict name -> uint64 counts [cardinality + 1], slot 0 = nulls (string columns)."""
    c = _cache(frame, "codes")
    todo = [n for n in names if n not in c]
    if todo:
        for n, h in zip(todo, engine.code_counts(frame, todo)):
            c[n] = h
    return {n: c[n] for n in names}


def mode_distinct(frame: ColumnFrame, names):
    """dict name -> (mode value | None, mode_rows | None, n_distinct) over non-null values.
    Ties: smallest value (numeric) / smallest UTF-8 string (categorical); the reference's
    choice is arbitrary (stats_generator.py:358)."""
    c = _cache(frame, "mode")
    todo = [n for n in names if n not in c]
    cat = [n for n in todo if frame.column(n).kind == "cat"]
    num = [n for n in todo if frame.column(n).kind == "num"]
    if cat:
        cc = code_counts(frame, cat)
        for n in cat:
            h = cc[n][1:]
            dic = frame.column(n).dictionary
            if len(dic) == 0 or h.sum() == 0:
                c[n] = (None, None, 0)
                continue
            mx = int(h.max())
            best = min((dic[i] for i in np.flatnonzero(h == mx)), key=lambda s: s.encode("utf-8"))
            c[n] = (best, mx, int(np.count_nonzero(h)))
    if num:
        # the sort leaves the keys fully ordered: read the summary() percentiles out of it as
        # well, so a full stats_generator run never needs a separate selection pass
        mom = moments(frame, num)
        qc = _cache(frame, "quantiles")
        rk = np.array([engine.quantile_ranks(int(mom[n]["n_valid"]), SUMMARY_PROBS, SUMMARY_EPS) for n in num], dtype=np.int64)
        # ... and the HyperLogLog++ registers of the default rsd (p = 9) come out of the same pass: the run summaries hash one
        # key per distinct value, so a following measures_of_cardinality() needs no pass of its own
        fused_p = DEFAULT_HLL_P if (getattr(engine, "FUSED_HLL", False) and not getattr(frame, "is_partitioned", False)) else None
        if fused_p is not None:
            res, vals, regs = engine.sort_mode_distinct(frame, num, rk, hll_p=fused_p)
            if regs is not None:
                hc = _cache(frame, ("hll", fused_p))
                for n, r in zip(num, engine.hll_estimates_from_register_rows(regs, fused_p)):
                    hc.setdefault(n, r)
        else:
            res, vals = engine.sort_mode_distinct(frame, num, rk)
        rkl, vl = rk.tolist(), vals.tolist()
        for i, n in enumerate(num):
            c[n] = res[i]
            d = qc.get(n)
            if d is None:
                d = qc[n] = {}
            d.update(zip(rkl[i], vl[i]))
            d.pop(0, None)                     # rank 0 = "skip" (empty column)
    return {n: c[n] for n in names}


def hll(frame: ColumnFrame, names, rsd):
    """dict name -> (estimate, in_bias_band) of Spark's approx_count_distinct(col, rsd)."""
    rsd = 0.05 if rsd is None else rsd
    p = int(math.ceil(2.0 * math.log(1.106 / rsd) / math.log(2.0)))
    c = _cache(frame, ("hll", p))
    todo = [n for n in names if n not in c]
    if todo:
        for n, r in zip(todo, engine.hll_estimates(frame, todo, p)):
            c[n] = r
    return {n: c[n] for n in names}


def prefetch(frame: ColumnFrame, names=None, want=("moments", "mode", "hll"), rsd=None, group: int = 5):
    """Pipelined warm-up of the per-frame cache for a host-resident frame: every column's H2D
    copy is enqueued up front on a dedicated copy stream, and column group g is processed
    (moments -> sort-based mode/distinct/percentiles -> HLL++) while groups g+1.. are still in
    flight over PCIe.  The stats functions called afterwards hit the cache.  On a frame that is
    already on the device this is just the batched passes."""
    if getattr(frame, "is_partitioned", False):
        return frame  # chunks are uploaded (and freed) pass by pass
    torch = _lib.require_cuda()
    names = [n for n in (names or frame.columns) if frame.column(n).kind != "other"]
    from .frame import side_stream
    copy = side_stream()
    for n in names:
        frame.column(n).upload_async(copy)
    for g0 in range(0, len(names), group):
        grp = names[g0:g0 + group]
        if "moments" in want:
            moments(frame, grp)
        if "mode" in want:
            mode_distinct(frame, grp)
        if "hll" in want:
            hll(frame, grp, rsd)
    return frame
