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
    import warnings
    from . import api as O
    warnings.filterwarnings("ignore", message="No .* Computation", category=UserWarning)   # string-only column groups
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
    """Wall seconds of the streamed step over (table, target), one column per task over `workers` processes."""
    global _TABLE, _TARGET
    _TABLE, _TARGET = table, target
    workers = max(1, min(workers or os.cpu_count() or 1, len(table.column_names)))
    tasks = [[n] for n in _cost_order(table, table.column_names)]
    with mp.get_context("fork").Pool(workers, initializer=_all_cpus) as pool:
        pool.map(_stats_group, tasks[:workers], chunksize=1)  # warm the workers (imports), not timed
        t0 = time.perf_counter()
        for _ in pool.imap_unordered(_stream_group, tasks, chunksize=1):
            pass
        return time.perf_counter() - t0, workers


def _all_cpus():
    """Worker initializer: bench.py binds ITS process to the CPUs of the GPU's NUMA node (for the pinned buffers); the CPU
    arm should use every host core the container allows, so the children widen their affinity again (best effort)."""
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except (AttributeError, OSError, ValueError):
        pass


def _cost_order(table, names):
    """Longest first (string columns cost about 3x a numeric one in the oracle): with one column per task and the idle
    worker taking the next one, no worker is left holding two slow columns at the end."""
    import pyarrow as pa
    slow = [n for n in names if pa.types.is_string(table.schema.field(n).type) or pa.types.is_large_string(table.schema.field(n).type)
            or pa.types.is_dictionary(table.schema.field(n).type)]
    slow_set = set(slow)
    return slow + [n for n in names if n not in slow_set]


def time_stats_generator(table, workers=None, target=None):
    """Wall seconds of the 6 measures_of_* functions (+ drift when `target` is given) over
    `table`: one column per task, handed to `workers` processes as they become free (fork: the table is shared, not pickled)."""
    global _TABLE, _TARGET
    _TABLE, _TARGET = table, target
    workers = max(1, min(workers or os.cpu_count() or 1, len(table.column_names)))
    tasks = [[n] for n in _cost_order(table, table.column_names)]
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_all_cpus) as pool:
        pool.map(_stats_group, [t for t in tasks[:workers]], chunksize=1)  # warm the workers (imports), not timed
        t0 = time.perf_counter()
        for _ in pool.imap_unordered(_stats_group, tasks, chunksize=1):
            pass
        t_stats = time.perf_counter() - t0
        t_drift = None
        if target is not None:
            t0 = time.perf_counter()
            for _ in pool.imap_unordered(_drift_group, tasks, chunksize=1):
                pass
            t_drift = time.perf_counter() - t0
    return t_stats, t_drift, workers


# ---- parity leg of bench.py: the oracle's answers for a few columns of the bench frame at full length ---------------

_MAKE = None   # (column id, seed, shifted) -> one-column pyarrow Table; set by oracle_columns before the fork


def _oracle_column(job):
    c, with_drift = job
    import tempfile
    from . import api as O
    t = _MAKE(c, 42, False)
    name = t.column_names[0]
    out = {}
    for fn in ("measures_of_counts", "measures_of_centralTendency", "measures_of_cardinality", "measures_of_dispersion",
               "measures_of_percentiles", "measures_of_shape"):
        df = getattr(O, fn)(t)
        rec = df[df["attribute"] == name]
        out[fn] = None if len(rec) == 0 else {k: v for k, v in rec.iloc[0].to_dict().items() if k != "attribute"}
    if with_drift:
        tt = _MAKE(c, 43, True)
        with tempfile.TemporaryDirectory() as d:
            r = O.statistics(tt, t, method_type="all", use_sampling=False, source_path=d)
        out["drift"] = {m: float(r[m].iloc[0]) for m in ("PSI", "HD", "JSD", "KS")}
    return name, out


def oracle_columns(make_table, column_ids, drift_ids=()):
    """dict column name -> {function name -> {field -> value}, "drift" -> metrics}: the oracle on the tables
    `make_table(column id, seed, shifted)` builds (bench.py: the NumPy twin of the device generator), one process per column."""
    global _MAKE
    _MAKE = make_table
    jobs = [(c, c in set(drift_ids)) for c in column_ids]
    with mp.get_context("fork").Pool(max(1, len(jobs))) as pool:
        return dict(pool.map(_oracle_column, jobs))
