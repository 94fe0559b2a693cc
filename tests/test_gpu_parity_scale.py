"""Product vs ORACLE at the sizes the performance numbers are quoted on (BASELINE.json configs[1] / [2]):

* the device generator (csrc/synth.cu) is bit-identical to its NumPy twin (anovos_b200/synth.py), so the oracle and the
  GPU see the same values at any size;
* 10 M rows x 12 columns (all four synthetic families and null rates, float64 / int32 / int64 / string columns):
  every measures_of_* function, the mode / distinct / HLL++ paths, attribute_binning ids (bit-exact) and
  drift statistics(method_type="all") (<= 1e-6, north_star) against the oracle through the public API;
* 100 M rows x 2 columns (the c3 tile geometry: 100 M-row columns, 2^27-element tiles of the sort path).

Tolerances (north_star): counts, extrema, order statistics, modes, distinct counts, HLL++ estimates and bin ids
bit-exact; mean / stddev / skewness / kurtosis <= 1e-6 relative; PSI / HD / JSD / KS <= 1e-6 (met at 1e-9)."""
import math

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

from golden_util import assert_frames_match
from oracle import api as O
from oracle import spark_semantics as S

N10 = 10_000_000


def _bits(valid):
    b = np.packbits(valid, bitorder="little")
    return np.concatenate([b, np.zeros((-len(b)) % 4, np.uint8)]).view(np.int32)


@pytest.mark.parametrize("rows,row0", [(1_000_037, 0), (300_001, 65_536)])
def test_device_generator_equals_numpy_twin(rows, row0):
    """"(seed, column, row) so CPU baseline and GPU see identical values"."""
    from anovos_b200 import synth
    fr = synth.device_frame(rows, 8, seed=42, row0=row0)
    sh = synth.device_frame(rows, 8, seed=43, shifted=True, row0=row0)
    for frame, seed, shifted in ((fr, 42, False), (sh, 43, True)):
        for c, name in enumerate(frame.columns):
            d, v = frame.column(name).device()
            x, valid = synth.host_column(rows, c, seed, shifted, row0=row0)
            assert np.array_equal(d.cpu().numpy().view(np.uint32), x.view(np.uint32)), (name, seed)
            if v is None:
                assert valid.all()
            else:
                assert np.array_equal(v.cpu().numpy(), _bits(valid)), (name, seed)
    cat = synth.device_frame(rows, 16, seed=42, cat_every=4, row0=row0)
    for c in (4, 5, 6, 12):   # numeric columns of a mixed frame: family / null rate cycle over the numeric ordinal
        d, v = cat.column(cat.columns[c]).device()
        x, valid = synth.host_column(rows, c, 42, row0=row0, cat_every=4)
        assert np.array_equal(d.cpu().numpy().view(np.uint32), x.view(np.uint32)), c
        assert (v is None and valid.all()) or np.array_equal(v.cpu().numpy(), _bits(valid)), c
    for c in (3, 7, 11, 15):
        name = cat.columns[c]
        d, v = cat.column(name).device()
        codes, valid, dic = synth.host_codes(rows, c, 4, 42, row0=row0)
        assert cat.column(name).dictionary == dic
        assert np.array_equal(d.cpu().numpy(), codes), name
        if v is None:
            assert valid.all()
        else:
            assert np.array_equal(v.cpu().numpy(), _bits(valid)), name


def _table(n, seed, shifted):
    """8 synthetic float32 columns (bit-identical to the bench generator) + float64 / int32 / int64 / string."""
    from anovos_b200 import synth
    t = synth.host_table(n, 8, seed=seed, shifted=shifted)
    rng = np.random.default_rng(seed)
    k = 1.25 if shifted else 1.0

    def nulls(r):
        return rng.random(n) < r
    t = t.append_column("d_f64", pa.array(np.round(rng.normal(-1e6, 250.0 * k, n), 1), mask=nulls(0.05)))
    t = t.append_column("i_i32", pa.array(rng.integers(0, 110 if shifted else 100, n).astype(np.int32), mask=nulls(0.1)))
    t = t.append_column("l_i64", pa.array(rng.integers(-5000, 5000, n).astype(np.int64) * 1_000_003))
    s = synth.host_table(n, 1, seed=seed, first_col=15, cat_every=4, prefix="s")   # Zipf over 10 000 keys, 30 % nulls
    return t.append_column("s_cat", s.column(0))


@pytest.fixture(scope="module")
def big():
    from anovos_b200.frame import ColumnFrame
    t = _table(N10, 42, False)
    return t, ColumnFrame.from_arrow(t)


@pytest.fixture(scope="module")
def big_target():
    from anovos_b200.frame import ColumnFrame
    t = _table(N10, 43, True)
    return t, ColumnFrame.from_arrow(t)


def test_stats_generator_vs_oracle_10m(big):
    import anovos.data_analyzer.stats_generator as sg
    from anovos_b200 import profile
    t, fr = big
    for fn, kw in (("measures_of_counts", {}), ("measures_of_centralTendency", {}),
                   ("measures_of_cardinality", {"use_approx_unique_count": False}), ("measures_of_cardinality", {}),
                   ("measures_of_dispersion", {}), ("measures_of_percentiles", {}), ("measures_of_shape", {}),
                   ("missingCount_computation", {}), ("nonzeroCount_computation", {}), ("mode_computation", {}),
                   ("uniqueCount_computation", {})):
        assert_frames_match(getattr(sg, fn)(None, fr, **kw).toPandas(), getattr(O, fn)(t, **kw))
    # unrounded moments through the API's own cache vs the oracle's raw outputs: <= 1e-6 relative
    num = [c for c in fr.columns if fr.column(c).kind == "num"]
    mom = profile.moments(fr, num)
    ct = O.measures_of_centralTendency(t, raw=True).set_index("attribute")
    dp = O.measures_of_dispersion(t, raw=True).set_index("attribute")
    shp = O.measures_of_shape(t, raw=True).set_index("attribute")
    for c in num:
        m = mom[c]
        n, m2, m3, m4 = float(m["n_valid"]), float(m["m2"]), float(m["m3"]), float(m["m4"])
        got = {"mean": float(m["mean"]), "stddev": math.sqrt(m2 / (n - 1)), "skewness": math.sqrt(n) * m3 / m2 ** 1.5,
               "kurtosis": n * m4 / (m2 * m2) - 3.0}
        exp = {"mean": ct.loc[c, "mean"], "stddev": dp.loc[c, "stddev"], "skewness": shp.loc[c, "skewness"],
               "kurtosis": shp.loc[c, "kurtosis"]}
        for k in got:
            scale = max(abs(exp[k]), 1e-3 if k in ("skewness", "kurtosis") else 0.0)   # skew / kurt near 0: absolute
            assert abs(got[k] - exp[k]) <= 1e-6 * scale, (c, k, got[k], exp[k])


@pytest.mark.parametrize("method,bins", [("equal_range", 10), ("equal_frequency", 10), ("equal_range", 64)])
def test_attribute_binning_ids_vs_oracle_10m(big, method, bins):
    import anovos.data_transformer.transformers as tr
    t, fr = big
    cols = ["c0000", "c0001", "c0003", "c0006", "d_f64", "i_i32", "l_i64"]
    got = tr.attribute_binning(None, fr, list_of_cols=cols, method_type=method, bin_size=bins)
    exp = O.attribute_binning(t, list_of_cols=cols, method_type=method, bin_size=bins)
    for c in cols:
        ids = got.column(c).device()[0].cpu().numpy()
        assert np.array_equal(ids, np.asarray(exp.column(c).combine_chunks().fill_null(0))), (c, method)   # bit-exact


@pytest.mark.parametrize("method,bins", [("equal_range", 10), ("equal_frequency", 10)])
def test_drift_statistics_vs_oracle_10m(big, big_target, tmp_path, method, bins):
    import anovos.drift_stability.drift_detector as dd
    (ts, fs), (tt, ft) = big, big_target
    kw = dict(method_type="all", bin_method=method, bin_size=bins, use_sampling=False, threshold=0.1)
    got = dd.statistics(None, ft, fs, source_path=str(tmp_path / "g"), **kw).toPandas()
    exp = O.statistics(tt, ts, source_path=str(tmp_path / "o"), **kw)
    assert got["attribute"].tolist() == exp["attribute"].tolist()
    for m in ("PSI", "HD", "JSD", "KS"):
        g, e = got[m].values.astype(float), np.array([0.0 if v is None else v for v in exp[m].tolist()], float)
        assert np.allclose(g, e, rtol=1e-6, atol=1e-12), (m, g, e)      # north_star tolerance
        assert np.allclose(g, e, rtol=1e-9, atol=1e-12), (m, g, e)      # what the integer histograms actually give
    assert got["flagged"].tolist() == exp["flagged"].tolist()
    assert 0 < sum(got["flagged"]) < len(got)


def test_c3_geometry_100m_rows_vs_oracle(tmp_path):
    """Two columns of the c3 workload at its real length (100 M rows: a dense normal column and the zero-inflated,
    30 %-null one), generated on the device, against the oracle on the NumPy twin."""
    import anovos.data_analyzer.stats_generator as sg
    import anovos.drift_stability.drift_detector as dd
    from anovos_b200 import synth
    from anovos_b200.frame import ColumnFrame
    n = 100_000_000
    ids = [0, 3]
    whole = synth.device_frame(n, 4, seed=42)
    fr = whole.select([whole.columns[i] for i in ids])
    t = synth.host_table(n, 4, seed=42, columns=ids)
    assert t.column_names == fr.columns
    for fn, kw in (("measures_of_counts", {}), ("measures_of_centralTendency", {}),
                   ("measures_of_cardinality", {"use_approx_unique_count": False}), ("measures_of_cardinality", {}),
                   ("measures_of_dispersion", {}), ("measures_of_percentiles", {}), ("measures_of_shape", {})):
        assert_frames_match(getattr(sg, fn)(None, fr, **kw).toPandas(), getattr(O, fn)(t, **kw))
    whole_t = synth.device_frame(n, 4, seed=43, shifted=True)
    ft = whole_t.select([whole_t.columns[i] for i in ids])
    tt = synth.host_table(n, 4, seed=43, shifted=True, columns=ids)
    kw = dict(method_type="all", use_sampling=False)
    got = dd.statistics(None, ft, fr, source_path=str(tmp_path / "g"), **kw).toPandas()
    exp = O.statistics(tt, t, source_path=str(tmp_path / "o"), **kw)
    for m in ("PSI", "HD", "JSD", "KS"):
        assert np.allclose(got[m].values.astype(float), np.asarray(exp[m], float), rtol=1e-9, atol=1e-12), m
    assert got["flagged"].tolist() == exp["flagged"].tolist()


def test_batched_offsets_beyond_2_31_bytes():
    """Six 100 M-row columns in ONE call: key buffers, bin-id output and tile tables span > 2^31 bytes.  The batched
    results must equal the per-column calls (which the 100 M-row oracle test pins)."""
    from anovos_b200 import engine, synth
    n = 100_000_000
    fr = synth.device_frame(n, 6, seed=42)
    names = fr.columns
    mom = engine.moments(fr, names)
    rk = np.array([[1, int(m["n_valid"]) // 2, int(m["n_valid"])] for m in mom], np.int64)
    res, vals = engine.sort_mode_distinct(fr, names, rk)
    for i in (0, 5):
        r1, v1 = engine.sort_mode_distinct(fr, [names[i]], rk[i:i + 1])
        assert r1[0] == res[i] and np.array_equal(v1[0], vals[i])
        assert vals[i][0] == mom["min"][i] and vals[i][2] == mom["max"][i]
    cuts, lohi = [], []
    for m in mom:
        w = (float(m["max"]) - float(m["min"])) / 10
        cuts.append([float(m["min"]) + j * w for j in range(1, 10)])
        lohi.append((float(m["min"]), float(m["max"])))
    model = engine.BinModel(fr, names, cuts, lohi)
    ids = engine.bin_assign(fr, model)
    h = engine.histogram(fr, model)
    for i in (0, 5):
        cnt = np.bincount(ids[i].cpu().numpy(), minlength=11)
        assert np.array_equal(cnt.astype(np.uint64), h[i][:11]), names[i]
