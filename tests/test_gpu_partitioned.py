"""GPU tests of the row-partitioned path (SURVEY.md 8e row-sharded variant / C4-C5 streaming):
every merged pass must reproduce the resident-frame result - counts, bin indices, percentiles and
HLL estimates bit-exactly, FP moments and drift within 1e-9 relative (merge order differs) - and
the income golden values must survive chunking.  The 2-rank tests run two processes on ONE GPU
with a gloo group (the collectives are the same calls NCCL serves on a multi-GPU box)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mixed(rows, seed=42, row0=0, cols=8):
    from anovos_b200 import synth
    return synth.device_frame(rows, cols, seed=seed, cat_every=4, row0=row0)


def _host(col):
    d, v = col.device()
    return d.cpu().numpy(), (None if v is None else v.cpu().numpy())


def test_synth_chunk_is_a_slice_of_the_whole():
    whole = _mixed(100_000)
    part = _mixed(100_000 - 32_000, row0=32_000)
    for n in whole.columns:
        d, v = _host(whole.column(n))
        pd_, pv = _host(part.column(n))
        assert np.array_equal(d[32_000:], pd_), n
        if v is not None:
            assert np.array_equal(v[1000:], pv), n


def _close(a, b, rtol=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.allclose(a, b, rtol=rtol, atol=0, equal_nan=True)


def _same_table(a, b):
    """Cell-wise equality of two result tables (NaN == NaN; ignores int/float/object dtype differences
    introduced by shipping a table between processes as a dict)."""
    import pandas as pd
    assert list(a.columns) == list(b.columns) and len(a) == len(b)
    a, b = a.sort_values("attribute").reset_index(drop=True), b.sort_values("attribute").reset_index(drop=True)
    for c in a.columns:
        for x, y in zip(a[c].tolist(), b[c].tolist()):
            assert (pd.isna(x) and pd.isna(y)) or x == y, (c, x, y)


def _moments_close(a, b):
    """mean / M2 / M4 within 1e-9 relative; M3 is a cancelling sum, so it is held to 1e-9 of its natural
    scale M2^1.5 / sqrt(n) (= skewness within 1e-9 absolute)."""
    for f in ("mean", "m2", "m4"):
        assert _close(a[f], b[f]), f
    with np.errstate(all="ignore"):
        scale = a["m2"] ** 1.5 / np.sqrt(np.maximum(a["n_valid"], 1))
    assert np.all(np.abs(a["m3"] - b["m3"]) <= 1e-9 * scale + 1e-9 * np.abs(a["m3"])), "m3"


def _check_against_whole(whole, parts, tmp_path, exact_mode=False):
    """whole: resident ColumnFrame; parts: PartitionedFrame over the same rows."""
    import anovos.data_analyzer.stats_generator as sg
    import anovos.data_transformer.transformers as tr
    import anovos.drift_stability.drift_detector as dd
    from anovos_b200 import engine
    from anovos_b200.shared.utils import attributeType_segregation
    num, cat, _ = attributeType_segregation(whole)
    assert parts.count() == whole.count() and parts.columns == whole.columns
    mw, mp_ = engine.moments(whole, num + cat), engine.moments(parts, num + cat)
    for f in ("n_valid", "n_nonzero", "min", "max"):
        assert np.array_equal(mw[f], mp_[f], equal_nan=True), f
    _moments_close(mw, mp_)
    for fn in ("global_summary", "measures_of_counts", "measures_of_percentiles", "measures_of_cardinality",
               "missingCount_computation", "nonzeroCount_computation"):
        a, b = getattr(sg, fn)(None, whole).toPandas(), getattr(sg, fn)(None, parts).toPandas()
        assert a.equals(b), (fn, a, b)
    for fn in ("measures_of_dispersion", "measures_of_shape"):
        a, b = getattr(sg, fn)(None, whole).toPandas(), getattr(sg, fn)(None, parts).toPandas()
        assert list(a["attribute"]) == list(b["attribute"])
        va, vb = a.drop(columns="attribute").to_numpy(float), b.drop(columns="attribute").to_numpy(float)
        assert np.allclose(va, vb, rtol=1e-9, atol=1.01e-4, equal_nan=True), fn   # values are rounded to 4 decimals
    if cat:  # the mode of string columns merges (code histograms add)
        a, b = sg.mode_computation(None, whole, cat).toPandas(), sg.mode_computation(None, parts, cat).toPandas()
        assert a.equals(b)
    # numeric mode / exact distinct need whole columns: local chunks are concatenated on the device when they fit
    a, b = sg.measures_of_centralTendency(None, whole).toPandas(), sg.measures_of_centralTendency(None, parts).toPandas()
    _same_table(a, b)
    a, b = sg.uniqueCount_computation(None, whole).toPandas(), sg.uniqueCount_computation(None, parts).toPandas()
    _same_table(a, b)
    # binning: the chunks of the binned frame are the row slices of the resident result
    for method in ("equal_range", "equal_frequency"):
        bw = tr.attribute_binning(None, whole, num, method_type=method, bin_size=7)
        bp = tr.attribute_binning(None, parts, num, method_type=method, bin_size=7)
        assert bp.columns == bw.columns and bp.count() == bw.count()
        r0 = 0
        for ch in bp.chunks():
            for n in num:
                assert np.array_equal(_host(ch.column(n))[0], _host(bw.column(n))[0][r0:r0 + ch.n_rows]), (method, n)
            r0 += ch.n_rows
        assert r0 == whole.n_rows
    return num, cat


@pytest.mark.parametrize("rows,chunk", [(100_000, 32_768), (70_001, 9_984), (5_000, 1 << 20)])
def test_chunked_frame_matches_resident(rows, chunk, tmp_path):
    from anovos_b200.partitioned import PartitionedFrame
    import anovos.drift_stability.drift_detector as dd
    whole = _mixed(rows)
    parts = PartitionedFrame.from_frame(whole, chunk)
    assert parts.n_chunks == -(-rows // (chunk // 32 * 32))
    _check_against_whole(whole, parts, tmp_path)
    tgt_w = _mixed(rows - 777, seed=43)
    tgt_p = PartitionedFrame.from_frame(tgt_w, chunk)
    kw = dict(method_type="all", use_sampling=False, bin_size=10)
    a = dd.statistics(None, tgt_w, whole, source_path=str(tmp_path / "a"), **kw).toPandas()
    b = dd.statistics(None, tgt_p, parts, source_path=str(tmp_path / "b"), **kw).toPandas()
    assert list(a["attribute"]) == list(b["attribute"]) and list(a["flagged"]) == list(b["flagged"])
    for m in ("PSI", "HD", "JSD", "KS"):
        assert _close(a[m], b[m], 1e-9), m
    # mixed: resident source, partitioned target + a source model saved by the partitioned run
    c = dd.statistics(None, tgt_p, None, pre_existing_source=True, source_path=str(tmp_path / "b"), **kw).toPandas()
    for m in ("PSI", "HD", "JSD", "KS"):
        assert _close(a[m], c[m], 1e-9), m


def test_chunked_host_table_and_golden(income, nb_stats, tmp_path):
    """The income dataset, uploaded chunk by chunk (H2D of chunk i+1 in flight), against the
    resident frame AND the reference's stored Spark outputs."""
    import pandas as pd
    from golden_util import frame_by_attr, shown_close, table_by_attr
    import anovos.data_analyzer.stats_generator as sg
    from anovos_b200.frame import ColumnFrame
    from anovos_b200.partitioned import PartitionedFrame
    whole = ColumnFrame.from_arrow(income)
    parts = PartitionedFrame.from_frame(income, 4096)
    assert parts.n_chunks == 8
    _check_against_whole(whole, parts, tmp_path)
    for df, t, cols in ((sg.measures_of_shape(None, parts), nb_stats[39], ["skewness", "kurtosis"]),
                        (sg.measures_of_counts(None, parts), nb_stats[11], ["fill_count", "missing_count", "nonzero_count"]),
                        (sg.measures_of_dispersion(None, parts), nb_stats[31], ["stddev", "variance", "cov", "range"]),
                        (sg.measures_of_cardinality(None, parts), nb_stats[23], ["unique_values", "IDness"])):
        got, exp = frame_by_attr(df.toPandas()), table_by_attr(t)
        assert set(got) == set(exp)
        bad = [(a, c, got[a][c], row[c]) for a, row in exp.items() for c in cols
               if not shown_close(None if pd.isna(got[a][c]) else got[a][c], row[c])]
        assert not bad, bad


def test_select_split_equals_monolithic():
    from anovos_b200 import engine
    from anovos_b200.partitioned import PartitionedFrame
    fr = _mixed(50_000, cols=6)
    num = [n for n in fr.columns if fr.column(n).kind == "num"]
    mom = engine.moments(fr, num)
    rk = np.array([[1, max(1, int(m["n_valid"]) // 3), max(1, int(m["n_valid"]))] for m in mom], dtype=np.int64)
    a = engine.select_ranks(fr, num, rk)
    b = engine.select_ranks(PartitionedFrame.from_frames([fr]), num, rk)
    c = engine.select_ranks(PartitionedFrame.from_frame(fr, 4096), num, rk)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.array_equal(a[:, 0], mom["min"]) and np.array_equal(a[:, 2], mom["max"])


# ---- two ranks, row slabs ------------------------------------------------------------------------------

def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


ROWS, SLAB0 = 90_000, 40_000 // 32 * 32


def _rank_worker(rank, world, port, ret):
    import torch.distributed as dist
    import anovos.data_analyzer.stats_generator as sg
    import anovos.drift_stability.drift_detector as dd
    from anovos_b200 import engine, parallel
    from anovos_b200.partitioned import PartitionedFrame, repartition_to_columns
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r0, r1 = (0, SLAB0) if rank == 0 else (SLAB0, ROWS)
    slab = _mixed(r1 - r0, row0=r0)
    tslab = _mixed(r1 - r0, seed=43, row0=r0)
    # (a) row slabs merged in place: chunked locally AND across ranks
    parts = PartitionedFrame.from_frame(slab, 16_384, group=True)
    tparts = PartitionedFrame.from_frame(tslab, 16_384, group=True)
    num = [n for n in slab.columns if slab.column(n).kind == "num"]
    out = {"count": parts.count(), "mom": engine.moments(parts, slab.columns).tolist()}
    for fn in ("measures_of_counts", "measures_of_percentiles", "measures_of_cardinality", "measures_of_shape"):
        out[fn] = getattr(sg, fn)(None, parts).toPandas().to_dict("list")
    # exact mode / distinct on row slabs: the slabs are exchanged into column blocks under the hood and the per-column
    # results all-gathered, so every rank holds the whole table
    out["central_slabs"] = sg.measures_of_centralTendency(None, parts).toPandas().to_dict("list")
    out["unique_slabs"] = sg.uniqueCount_computation(None, parts).toPandas().to_dict("list")
    out["drift"] = dd.statistics(None, tparts, parts, method_type="all", use_sampling=False,
                                 source_path="/tmp/anv_part_test_%d_%d" % (port, rank)).toPandas().to_dict("list")
    # (b) the exchange: row slabs -> whole columns of this rank's block, then the full column path (incl. exact mode)
    mine = repartition_to_columns(slab, True)
    assert mine.columns == parallel.shard_columns(slab.columns, rank, world) and mine.count() == ROWS
    out["central"] = sg.measures_of_centralTendency(None, mine).toPandas().to_dict("list")
    out["exact_unique"] = sg.uniqueCount_computation(None, mine).toPandas().to_dict("list")
    ret[rank] = out
    dist.destroy_process_group()


def test_two_ranks_row_slabs_match_single_frame(tmp_path):
    import torch.multiprocessing as mp
    import pandas as pd
    import anovos.data_analyzer.stats_generator as sg
    import anovos.drift_stability.drift_detector as dd
    from anovos_b200 import engine
    world, port = 2, _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_rank_worker, args=(world, port, ret), nprocs=world, join=True)
    a, b = ret[0], ret[1]
    whole, twhole = _mixed(ROWS), _mixed(ROWS, seed=43)
    for k in ("count", "mom", "measures_of_counts", "measures_of_percentiles", "measures_of_cardinality",
              "measures_of_shape", "drift"):
        assert repr(a[k]) == repr(b[k]), k                      # every rank takes the same decisions
    assert a["count"] == ROWS
    for k in ("central_slabs", "unique_slabs"):
        assert repr(a[k]) == repr(b[k]), k
    mw = engine.moments(whole, whole.columns)
    mp_ = np.array([tuple(r) for r in a["mom"]], dtype=engine._MOM_DT)
    for f in ("n_valid", "n_nonzero", "min", "max"):
        assert np.array_equal(mw[f], mp_[f], equal_nan=True), f
    _moments_close(mw, mp_)
    for fn in ("measures_of_counts", "measures_of_percentiles", "measures_of_cardinality"):
        assert getattr(sg, fn)(None, whole).toPandas().equals(pd.DataFrame(a[fn])), fn
    sh = sg.measures_of_shape(None, whole).toPandas()
    assert np.allclose(sh.drop(columns="attribute").to_numpy(float),
                       pd.DataFrame(a["measures_of_shape"]).drop(columns="attribute").to_numpy(float), atol=1.01e-4)
    dw = dd.statistics(None, twhole, whole, method_type="all", use_sampling=False, source_path=str(tmp_path)).toPandas()
    for m in ("PSI", "HD", "JSD", "KS"):
        assert _close(dw[m], a["drift"][m], 1e-9), m
    assert list(dw["flagged"]) == a["drift"]["flagged"]
    # after the exchange each rank owns whole columns: the concatenation is the single-frame result
    cw = sg.measures_of_centralTendency(None, whole).toPandas()
    _same_table(cw, pd.DataFrame(a["central_slabs"]))
    _same_table(sg.uniqueCount_computation(None, whole).toPandas(), pd.DataFrame(a["unique_slabs"]))
    cp = pd.concat([pd.DataFrame(a["central"]), pd.DataFrame(b["central"])], ignore_index=True)
    _same_table(cw, cp)
    uw = sg.uniqueCount_computation(None, whole).toPandas()
    up = pd.concat([pd.DataFrame(a["exact_unique"]), pd.DataFrame(b["exact_unique"])], ignore_index=True)
    _same_table(uw, up)


def test_nccl_row_slabs_on_two_gpus():
    """The same checks over NCCL (device-side all_reduce of the select histograms, grouped P2P exchange);
    needs two GPUs - skipped on a single-GPU box, where the gloo test above covers the logic."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(root, "scripts", "nccl_rowslab_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ROWSLAB-OK 2" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
