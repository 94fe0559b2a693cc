"""CPU timing of the oracle on the host cores (bench.py `cpu_baseline` and `--impl reference`).
TEST / BENCH INFRASTRUCTURE ONLY.  Spark is not available on the box, so the reference arm
is this NumPy restatement run with one process per column group on all host cores
("CPU restatement, not Spark" - BASELINE.md section 3)."""
from __future__ import annotations

import multiprocessing as mp
import os
import time

_TABLE = None
_TARGET = None


def _stats_group(names):
    from . import api as O
    t = _TABLE.select(names)
    O.measures_of_counts(t)
    O.measures_of_centralTendency(t)
    O.measures_of_cardinality(t)
    O.measures_of_dispersion(t)
    O.measures_of_percentiles(t)
    O.measures_of_shape(t)
    return len(names)


def _drift_group(names):
    import tempfile
    from . import api as O
    with tempfile.TemporaryDirectory() as d:
        O.statistics(_TARGET.select(names), _TABLE.select(names), method_type="all", use_sampling=False,
                     source_path=d)
    return len(names)


def _stream_group(names):
    """The streamed bench step (bench.py c4/c5): drift statistics(all) + counts and shape of both frames."""
    from . import api as O
    _drift_group(names)
    for t in (_TABLE.select(names), _TARGET.select(names)):
        O.measures_of_counts(t)
        O.measures_of_shape(t)
    return len(names)


def time_stream_step(table, target, workers=None):
    """Wall seconds of the streamed step over (table, target), columns spread over `workers` processes."""
    global _TABLE, _TARGET
    _TABLE, _TARGET = table, target
    workers = workers or os.cpu_count() or 1
    groups = _split(table.column_names, workers)
    with mp.get_context("fork").Pool(len(groups)) as pool:
        pool.map(_stats_group, [g[:1] for g in groups])  # warm the workers (imports), not timed
        t0 = time.perf_counter()
        pool.map(_stream_group, groups)
        return time.perf_counter() - t0, len(groups)


def _split(names, k):
    k = max(1, min(k, len(names)))
    return [names[i::k] for i in range(k)]


def time_stats_generator(table, workers=None, target=None):
    """Wall seconds of the 6 measures_of_* functions (+ drift when `target` is given) over
    `table`, columns spread over `workers` processes (fork: the table is shared, not pickled)."""
    global _TABLE, _TARGET
    _TABLE, _TARGET = table, target
    workers = workers or os.cpu_count() or 1
    groups = _split(table.column_names, workers)
    ctx = mp.get_context("fork")
    with ctx.Pool(len(groups)) as pool:
        pool.map(_stats_group, [g[:1] for g in groups])  # warm the workers (imports), not timed
        t0 = time.perf_counter()
        pool.map(_stats_group, groups)
        t_stats = time.perf_counter() - t0
        t_drift = None
        if target is not None:
            t0 = time.perf_counter()
            pool.map(_drift_group, groups)
            t_drift = time.perf_counter() - t0
    return t_stats, t_drift, len(groups)
