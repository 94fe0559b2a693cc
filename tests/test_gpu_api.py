"""GPU parity tests through the reference-facing API (`anovos.*` module paths):
(1) the reference's own unit tests, ported with unchanged inputs and expected values
    (file:line relative to /root/reference/src/test/anovos),
(2) the stored Spark outputs of the reference notebooks on the income dataset,
(3) product vs oracle on seeded synthetic frames (nulls, ragged sizes, mixed dtypes)."""
import math
import os

import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

from golden_util import assert_frames_match, cell_value, frame_by_attr, shown_close, table_by_attr
from oracle import api as O
from oracle import spark_semantics as S


@pytest.fixture(scope="module")
def sg():
    import anovos.data_analyzer.stats_generator as m
    return m


@pytest.fixture(scope="module")
def dd():
    import anovos.drift_stability.drift_detector as m
    return m


@pytest.fixture(scope="module")
def tr():
    import anovos.data_transformer.transformers as m
    return m


def _df4():
    return O.table_from_rows([("27520a", 51, "HS-grad"), ("10a", 42, "Postgrad"), ("11a", 55, None),
                              ("1100b", 23, "HS-grad")], ["ifa", "age", "education"])


def _rec(res, attr):
    d = res.where({"attribute": attr}).toPandas().to_dict("list")
    return {k: v[0] for k, v in d.items()}


# ---- (1) data_analyzer/test_stats_generator.py ---------------------------------------------------

def test_missingCount_computation(sg):  # :29-65
    r = sg.missingCount_computation(None, _df4())
    assert r.count() == 3
    assert _rec(r, "education")["missing_count"] == 1 and _rec(r, "education")["missing_pct"] == 0.25


def test_uniqueCount_computation(sg):  # :68-184
    t = O.table_from_rows([("27520a", 51, 9000, "HS-grad"), ("10a", 42, 7000, "Postgrad"), ("11a", 35, None, None),
                           ("1100b", 23, 6000, "HS-grad")], ["ifa", "age", "income", "education"])
    for kw in ({}, {"compute_approx_unique_count": True}, {"compute_approx_unique_count": True, "rsd": 0.05},
               {"compute_approx_unique_count": True, "rsd": 0.2}):
        r = sg.uniqueCount_computation(None, t, **kw)
        assert r.count() == 4
        assert _rec(r, "education")["unique_values"] == 2 and _rec(r, "age")["unique_values"] == 4
        assert _rec(r, "income")["unique_values"] == 3
    with pytest.raises(ValueError):
        sg.uniqueCount_computation(None, t, compute_approx_unique_count=True, rsd=-1)


def test_mode_computation(sg):  # :187-235
    t = O.table_from_rows([("27520a", 51, "HS-grad"), ("10a", 42, "Postgrad"), ("11a", 55, None),
                           ("13a", 42, "HS-grad"), ("1100b", 23, "HS-grad")], ["ifa", "age", "education"])
    r = sg.mode_computation(None, t)
    assert r.count() == 3
    assert _rec(r, "education")["mode"] == "HS-grad" and _rec(r, "education")["mode_rows"] == 3
    assert _rec(r, "age")["mode"] == "42" and _rec(r, "age")["mode_rows"] == 2


def test_nonzeroCount_computation(sg):  # :238-289
    t = O.table_from_rows([("27520a", 51, 9000, "HS-grad"), ("10a", 0, 7000, "Postgrad"), ("11a", 35, None, None),
                           ("1100b", 23, 6000, "HS-grad")], ["ifa", "age", "income", "education"])
    r = sg.nonzeroCount_computation(None, t)
    assert r.count() == 2
    assert _rec(r, "age")["nonzero_count"] == 3 and _rec(r, "age")["nonzero_pct"] == 0.75
    assert _rec(r, "income")["nonzero_count"] == 3 and _rec(r, "income")["nonzero_pct"] == 0.75


def test_measures_of_centralTendency(sg):  # :292-339
    r = sg.measures_of_centralTendency(None, _df4())
    assert r.count() == 3
    a = _rec(r, "age")
    assert a["mean"] == 42.75 and a["median"] == 42.0
    e = _rec(r, "education")
    assert e["mode"] == "HS-grad" and e["mode_rows"] == 2 and e["mode_pct"] == 0.6667


def test_measures_of_cardinality(sg):  # :342-448
    for kw in ({}, {"use_approx_unique_count": False}):
        r = sg.measures_of_cardinality(None, _df4(), **kw)
        assert r.count() == 3
        assert _rec(r, "age")["unique_values"] == 4 and _rec(r, "age")["IDness"] == 1.0
        assert _rec(r, "education")["unique_values"] == 2 and _rec(r, "education")["IDness"] == 0.6667
        assert _rec(r, "ifa")["IDness"] == 1.0


def test_measures_of_dispersion(sg):  # :451-504
    a = _rec(sg.measures_of_dispersion(None, _df4()), "age")
    assert (a["stddev"], a["variance"], a["cov"], a["IQR"], a["range"]) == (14.2449, 202.9172, 0.3332, 28.0, 32.0)


def test_measures_of_counts(sg):  # :508-567
    r = sg.measures_of_counts(None, _df4())
    assert r.count() == 3
    a = _rec(r, "age")
    assert (a["fill_count"], a["fill_pct"], a["missing_count"], a["missing_pct"], a["nonzero_count"],
            a["nonzero_pct"]) == (4, 1.0, 0, 0.0, 4, 1.0)
    e = _rec(r, "education")
    assert (e["fill_count"], e["fill_pct"], e["missing_count"], e["missing_pct"]) == (3, 0.75, 1, 0.25)


def test_measures_of_shape(sg):  # :570-605
    a = _rec(sg.measures_of_shape(None, _df4()), "age")
    assert a["skewness"] == -0.7063 and a["kurtosis"] == -1.0646


def test_global_summary(sg):  # :608-661
    g = dict(sg.global_summary(None, _df4()).toPandas().values.tolist())
    assert g["rows_count"] == "4" and g["columns_count"] == "3" and g["numcols_count"] == "1"
    assert g["numcols_name"] == "age" and g["catcols_count"] == "2" and g["catcols_name"] == "ifa, education"


def test_measures_of_percentiles(sg):  # :664-779
    t = O.table_from_rows([("a", 51), ("b", 42), ("c", 55), ("d", 23), ("e", 46), ("f", 33)], ["ifa", "age"])
    a = _rec(sg.measures_of_percentiles(None, t), "age")
    assert a["min"] == 23.0 and a["max"] == 55.0 and a["50%"] == 42.0 and a["25%"] == 33.0 and a["99%"] == 55.0


def test_invalid_columns_raise(sg):
    with pytest.raises(TypeError):
        sg.measures_of_counts(None, _df4(), list_of_cols=["nope"])
    with pytest.raises(TypeError):
        sg.measures_of_dispersion(None, _df4(), list_of_cols=["education"])
    with pytest.warns(UserWarning):
        r = sg.measures_of_shape(None, _df4().select(["ifa", "education"]))
    assert r.count() == 0 and r.columns == ["attribute", "skewness", "kurtosis"]


# ---- (1) drift_stability/test_drift_detector.py:7-46, test_validations.py:13-20 -------------------------

def test_that_drift_statistics_can_be_calculated(dd, tmp_path):
    r = np.array([0.34, -1.76, 0.32, -0.39, -0.67, 0.61, 1.03, 0.93, -0.84, -0.31])
    tgt = pa.table({"A": r, "B": r})
    src = pa.table({"A": r, "B": r + 1})
    d = dd.statistics(None, tgt, src, method_type="all", source_path=str(tmp_path)).toPandas()
    e = dd.statistics(None, tgt, src, method_type="all", bin_method="equal_frequency", print_impact=True,
                      source_path=str(tmp_path)).toPandas()
    d.index, e.index = d["attribute"], e["attribute"]
    assert d.loc["A", "PSI":"KS"].tolist() == [0, 0, 0, 0]
    assert d.loc[["A", "B"], "flagged"].tolist() == [0, 1]
    np.testing.assert_almost_equal(d.loc["B", "PSI":"KS"].astype(float), [7.6776, 0.7091, 0.3704, 0.4999], 4)
    assert e.loc["A", "PSI":"KS"].tolist() == [0, 0, 0, 0]
    np.testing.assert_almost_equal(e.loc["B", "PSI":"KS"].astype(float), [3.0899, 0.4775, 0.1769, 0.4], 4)
    assert e.loc[["A", "B"], "flagged"].tolist() == [0, 1]
    assert list(d.columns) == ["attribute", "PSI", "HD", "JSD", "KS", "flagged"]


def test_drift_validations(dd, tmp_path):
    t = pa.table({"A": np.arange(5.0)})
    with pytest.raises(ValueError):
        dd.statistics(None, t, t, list_of_cols=[], source_path=str(tmp_path))
    with pytest.raises(ValueError):
        dd.statistics(None, t, t, list_of_cols=["A"], drop_cols=["A"], source_path=str(tmp_path))
    with pytest.raises(ValueError):
        dd.statistics(None, t, t, list_of_cols=["Z"], source_path=str(tmp_path))
    with pytest.raises(TypeError):
        dd.statistics(None, t, t, method_type="XYZ", source_path=str(tmp_path))
    with pytest.raises(TypeError):
        dd.statistics(None, t, t, list_of_cols=5, source_path=str(tmp_path))


# ---- (1) data_transformer/test_transformers.py:39-104 ------------------------------------------------

def test_attribute_binning(tr, income_part1, tmp_path):
    cols = ["age", "fnlwgt", "hours-per-week"]
    odf = tr.attribute_binning(None, income_part1, list_of_cols=cols, bin_size=20, model_path=str(tmp_path))
    exp = O.attribute_binning(income_part1, list_of_cols=cols, bin_size=20)
    for c in cols:
        d, v = odf.column(c).device()
        ids = d.cpu().numpy()
        assert ids[ids > 0].min() == 1 and ids.max() == 20
        e = exp.column(c).combine_chunks()
        assert np.array_equal(ids, np.asarray(e.fill_null(0)))          # bit-exact bin ids vs oracle
    d0, _ = odf.column("education-num").device()
    assert np.array_equal(d0.cpu().numpy(), np.asarray(income_part1.column("education-num").combine_chunks().fill_null(0)))
    app = tr.attribute_binning(None, income_part1, list_of_cols=cols, bin_size=20, output_mode="append")
    assert len(app.columns) == len(income_part1.column_names) + 3 and "age_binned" in app.columns
    # pre-existing model: same ids; unknown column -> IndexError("list index out of range")
    again = tr.attribute_binning(None, income_part1, list_of_cols=cols, bin_size=20, pre_existing_model=True,
                                 model_path=str(tmp_path))
    for c in cols:
        assert np.array_equal(again.column(c).device()[0].cpu().numpy(), odf.column(c).device()[0].cpu().numpy())
    with pytest.raises(IndexError):
        tr.attribute_binning(None, income_part1, list_of_cols=["capital-gain"], bin_size=20, pre_existing_model=True,
                             model_path=str(tmp_path))
    for bad in ({"bin_size": 1}, {"method_type": "foo"}, {"output_mode": "x"}, {"list_of_cols": ["workclass"]}):
        with pytest.raises(TypeError):
            tr.attribute_binning(None, income_part1, **{"list_of_cols": cols, **bad})
    # the saved model is the reference's parquet layout
    import pyarrow.parquet as pq
    m = pq.read_table(os.path.join(str(tmp_path), "attribute_binning"))
    assert m.column_names == ["attribute", "parameters"] and len(m.column("parameters")[0]) == 19
    # equal_frequency + categorical labels run and agree with the oracle on ids
    eq = tr.attribute_binning(None, income_part1, list_of_cols=cols, method_type="equal_frequency", bin_size=10)
    eo = O.attribute_binning(income_part1, list_of_cols=cols, method_type="equal_frequency", bin_size=10)
    for c in cols:
        assert np.array_equal(eq.column(c).device()[0].cpu().numpy(), np.asarray(eo.column(c).combine_chunks().fill_null(0)))
    lab = tr.attribute_binning(None, income_part1, list_of_cols=["age"], bin_size=5, bin_dtype="categorical")
    lo = O.attribute_binning(income_part1, list_of_cols=["age"], bin_size=5, bin_dtype="categorical")
    col = lab.column("age")
    codes, vwords = col.device()
    ok = np.unpackbits(vwords.cpu().numpy().view(np.uint8), bitorder="little")[:col.n_rows].astype(bool)
    got = [col.dictionary[i] if v else None for i, v in zip(codes.cpu().numpy()[:2000], ok[:2000])]
    assert got == lo.column("age").to_pylist()[:2000]


# ---- (2) notebook golden vectors through the product API -----------------------------------------------

def _check(df, t, cols, skip=(), only_present=False):
    got, exp = frame_by_attr(df.toPandas()), table_by_attr(t)
    if only_present:
        exp = {a: r for a, r in exp.items() if a in got}
    assert set(got) == set(exp)
    bad = [(a, c, got[a][c], row[c]) for a, row in exp.items() for c in cols
           if (a, c) not in skip and not shown_close(None if pd.isna(got[a][c]) else got[a][c], row[c])]
    assert not bad, bad


def test_nb_counts(sg, income, nb_stats):
    _check(sg.measures_of_counts(None, income), nb_stats[11],
           ["fill_count", "fill_pct", "missing_count", "missing_pct", "nonzero_count", "nonzero_pct"])


def test_nb_dispersion_shape(sg, income, nb_stats):
    _check(sg.measures_of_dispersion(None, income), nb_stats[31], ["stddev", "variance", "cov", "range"])
    _check(sg.measures_of_shape(None, income), nb_stats[39], ["skewness", "kurtosis"])


def test_nb_central_tendency(sg, income, nb_stats):
    df = sg.measures_of_centralTendency(None, income)
    _check(df, nb_stats[17], ["mean", "mode_rows", "mode_pct"])
    ora = frame_by_attr(O.measures_of_centralTendency(income))
    got = frame_by_attr(df.toPandas())
    for a in got:   # identical to the oracle incl. tie-breaks and exact-rank medians
        for c in ("median", "mode"):
            g, e = got[a][c], ora[a][c]
            assert (pd.isna(g) and (e is None or pd.isna(e))) or g == e, (a, c, g, e)


def test_nb_cardinality(sg, income, nb_stats):
    _check(sg.measures_of_cardinality(None, income, use_approx_unique_count=False), nb_stats[24], ["unique_values", "IDness"])
    _check(sg.measures_of_cardinality(None, income), nb_stats[23], ["unique_values", "IDness"])   # HLL++ p=9: 21/21
    df = sg.measures_of_cardinality(None, income, rsd=0.02)
    band = set(df.toPandas().attrs["hll_bias_band"])
    assert band == {"geohash", "logfnl", "longitude"}
    _check(df, nb_stats[25], ["unique_values", "IDness"], skip={(a, c) for a in band for c in ("unique_values", "IDness")})


def test_nb_percentiles_equal_oracle(sg, income):
    got = sg.measures_of_percentiles(None, income).toPandas()
    exp = O.measures_of_percentiles(income)
    assert got["attribute"].tolist() == exp["attribute"].tolist()
    assert np.array_equal(got.drop(columns="attribute").values.astype(float), exp.drop(columns="attribute").values.astype(float))


def test_nb_percentiles_exact_with_spark_partitions(sg, income, nb_stats):
    """The table tagged with Spark's partitioning of the notebook run (tests/golden/income_partitions.json): the product
    takes the per-partition sketch samples from the sort kernel, merges them like QuantileSummaries and reproduces
    ALL 81 stored summary() percentiles, every median and every IQR - and every other statistic as before."""
    import json
    from conftest import GOLDEN
    parts = json.load(open(os.path.join(GOLDEN, "income_partitions.json")))["rows_per_partition"]
    t = O.with_spark_partitions(income, parts)
    _check(sg.measures_of_percentiles(None, t), nb_stats[35], ["min"] + list(S.SUMMARY_PCTS) + ["max"])
    _check(sg.measures_of_centralTendency(None, t), nb_stats[17], ["mean", "median", "mode_rows", "mode_pct"])
    _check(sg.measures_of_dispersion(None, t), nb_stats[31], ["stddev", "variance", "cov", "IQR", "range"])
    _check(sg.measures_of_counts(None, t), nb_stats[11], ["fill_count", "missing_count", "nonzero_count"])
    _check(sg.measures_of_shape(None, t), nb_stats[39], ["skewness", "kurtosis"])
    _check(sg.measures_of_cardinality(None, t), nb_stats[23], ["unique_values", "IDness"])
    ora = frame_by_attr(O.measures_of_percentiles(t))
    got = frame_by_attr(sg.measures_of_percentiles(None, t).toPandas())
    for a in got:   # and identical to the oracle's full sketch, value for value
        for c in S.SUMMARY_PCTS:
            assert (pd.isna(got[a][c]) and ora[a][c] is None) or got[a][c] == ora[a][c], (a, c)


def test_read_dataset_as_spark_partitions(sg, income, tmp_path):
    """read_dataset(..., spark_cores=N) scans a CSV the way Spark's local[N] would: product == the oracle's full sketch on
    the same partitions (percentiles, medians, IQR), everything else == the unpartitioned result."""
    import pyarrow.csv as pacsv
    from anovos.data_ingest.data_ingest import read_dataset
    path = str(tmp_path / "income.csv")
    pacsv.write_csv(income, path)                                          # ~ 5.5 MB: two Hadoop splits at 4 MiB
    fr = read_dataset(None, path, "csv", {"header": "True", "inferSchema": "True", "spark_cores": 8})
    assert fr.spark_partitions and fr.n_chunks == 2 and fr.count() == income.num_rows and max(fr.chunk_rows) < 50000
    plain = read_dataset(None, path, "csv", {"header": "True", "inferSchema": "True"})
    t = O.with_spark_partitions(income, fr.chunk_rows)
    got, ora = frame_by_attr(sg.measures_of_percentiles(None, fr).toPandas()), frame_by_attr(O.measures_of_percentiles(t))
    for a in ora:
        for c in ["min", "max"] + list(S.SUMMARY_PCTS):
            assert got[a][c] == ora[a][c], (a, c, got[a][c], ora[a][c])
    d1, d2 = frame_by_attr(sg.measures_of_dispersion(None, fr).toPandas()), frame_by_attr(O.measures_of_dispersion(t))
    for a in d2:
        assert d1[a]["IQR"] == d2[a]["IQR"], a
    for fn in ("measures_of_counts", "measures_of_shape", "measures_of_cardinality"):
        a, b = getattr(sg, fn)(None, fr).toPandas(), getattr(sg, fn)(None, plain).toPandas()
        assert a.drop(columns="attribute").round(4).equals(b.drop(columns="attribute").round(4)), fn


def test_nb_drift(dd, income, income_source, nb_drift, tmp_path):
    df = dd.statistics(None, income, income_source, source_path=str(tmp_path)).toPandas()
    got, exp = frame_by_attr(df), table_by_attr(nb_drift[6])
    assert set(got) == set(exp)
    for a, row in exp.items():
        assert abs(got[a]["PSI"] - float(row["PSI"])) < 5e-7, (a, got[a]["PSI"], row["PSI"])
        assert got[a]["flagged"] == int(row["flagged"])
    # saved source model -> pre_existing_source=True reproduces the metrics without the source frame
    again = dd.statistics(None, income, None, pre_existing_source=True, source_path=str(tmp_path)).toPandas()
    assert np.allclose(again["PSI"].values, df["PSI"].values, rtol=0, atol=1e-12)
    d3 = dd.statistics(None, income, income_source, list_of_cols=["age", "education-num", "capital-gain", "hours-per-week"],
                       method_type=["JSD", "HD", "KS"], bin_size=100, source_path=str(tmp_path)).toPandas()
    assert list(d3.columns) == ["attribute", "HD", "JSD", "KS", "flagged"] and (d3[["HD", "JSD", "KS"]].abs().values < 1e-12).all()


# ---- (3) product vs oracle on synthetic frames ---------------------------------------------------------

def _synth(n, seed, shift=False):
    rng = np.random.default_rng(seed)
    def nulls(rate):
        return rng.random(n) < rate
    k = 1.3 if shift else 1.0
    cats = np.array(["cat_%05d" % i for i in range(300)], dtype=object)
    cols = {
        "x_norm": pa.array((rng.normal(50.0 + (2 if shift else 0), 4.0 * k, n)).astype(np.float32), mask=nulls(0.001)),
        "x_logn": pa.array(np.exp(rng.normal(0, 0.75, n)).astype(np.float32), mask=nulls(0.02)),
        "x_unif": pa.array(rng.uniform(-3, 9 * k, n).astype(np.float32)),
        "x_zero": pa.array(np.where(rng.random(n) < 0.7, 0.0, rng.exponential(2.0 * k, n)).astype(np.float32), mask=nulls(0.3)),
        "d_f64": pa.array(rng.normal(-1e6, 250.0, n), mask=nulls(0.05)),
        "i_i32": pa.array(rng.integers(0, 90 if shift else 80, n).astype(np.int32), mask=nulls(0.1)),
        "l_i64": pa.array(rng.integers(-5000, 5000, n).astype(np.int64)),
        "c_small": pa.array(cats[np.minimum(rng.geometric(0.4, n) - 1, 11)], mask=nulls(0.02)),
        "c_large": pa.array(cats[np.minimum((rng.pareto(1.2, n) * (1.5 if shift else 1)).astype(np.int64), 299)], mask=nulls(0.0005)),
        "all_null": pa.array(np.zeros(n, np.float32), mask=np.ones(n, bool)),
    }
    return pa.table(cols)


def _cmp_frames(got, exp):
    assert_frames_match(got.toPandas(), exp)


@pytest.mark.parametrize("n", [37, 20011, 400003])
def test_stats_generator_vs_oracle(sg, n):
    t = _synth(n, seed=n)
    fr_all = sg.measures_of_counts(None, t)
    _cmp_frames(fr_all, O.measures_of_counts(t))
    _cmp_frames(sg.measures_of_centralTendency(None, t), O.measures_of_centralTendency(t))
    _cmp_frames(sg.measures_of_cardinality(None, t, use_approx_unique_count=False), O.measures_of_cardinality(t, use_approx_unique_count=False))
    _cmp_frames(sg.measures_of_cardinality(None, t), O.measures_of_cardinality(t))
    _cmp_frames(sg.measures_of_dispersion(None, t), O.measures_of_dispersion(t))
    _cmp_frames(sg.measures_of_percentiles(None, t), O.measures_of_percentiles(t))
    _cmp_frames(sg.measures_of_shape(None, t), O.measures_of_shape(t))
    _cmp_frames(sg.missingCount_computation(None, t), O.missingCount_computation(t))
    _cmp_frames(sg.mode_computation(None, t), O.mode_computation(t))
    _cmp_frames(sg.nonzeroCount_computation(None, t), O.nonzeroCount_computation(t))
    assert sg.global_summary(None, t).toPandas().values.tolist() == O.global_summary(t).values.tolist()


@pytest.mark.parametrize("n,bin_method,bins", [(53, "equal_range", 10), (30011, "equal_range", 10), (30011, "equal_frequency", 10),
                                               (250007, "equal_range", 25), (250007, "equal_frequency", 7)])
def test_drift_vs_oracle(dd, tmp_path, n, bin_method, bins):
    src, tgt = _synth(n, seed=n), _synth(n + 17, seed=n + 1, shift=True)
    kw = dict(method_type="all", bin_method=bin_method, bin_size=bins, use_sampling=False, threshold=0.1)
    got = dd.statistics(None, tgt, src, source_path=str(tmp_path / "g"), **kw).toPandas()
    exp = O.statistics(tgt, src, source_path=str(tmp_path / "o"), **kw)
    assert got["attribute"].tolist() == exp["attribute"].tolist()
    for m in ("PSI", "HD", "JSD", "KS"):
        g, e = got[m].values.astype(float), np.array([0.0 if v is None else v for v in exp[m].tolist()], float)
        assert np.allclose(g, e, rtol=1e-9, atol=1e-12), (m, g, e)     # north_star: within 1e-6
    assert got["flagged"].tolist() == exp["flagged"].tolist()
    assert 0 < sum(got["flagged"]) < len(got)
    # the frequency_counts CSVs follow the reference layout [<col>, p]
    f = pd.read_csv(os.path.join(str(tmp_path / "g"), "drift_statistics", "frequency_counts", "x_norm", "part-00000.csv"))
    assert list(f.columns) == ["x_norm", "p"]
    again = dd.statistics(None, tgt, None, pre_existing_source=True, source_path=str(tmp_path / "g"), **kw).toPandas()
    for m in ("PSI", "HD", "JSD", "KS"):
        assert np.allclose(again[m].values.astype(float), got[m].values.astype(float), rtol=1e-9, atol=1e-12), m


def test_empty_and_all_null_frames(sg, dd, tr, tmp_path):
    """Edge frames (parity unpinned: no reference test has them): zero rows, all-null columns, a single row.
    The product must not crash and must agree with the oracle cell for cell (nulls included)."""
    frames = {
        "empty": pa.table({"a": pa.array([], pa.float32()), "b": pa.array([], pa.int64()), "s": pa.array([], pa.string())}),
        "all_null": pa.table({"a": pa.array([None] * 5, pa.float64()), "b": pa.array([1, 2, None, 4, 5], pa.int32()),
                              "s": pa.array([None] * 5, pa.string())}),
        "one_row": pa.table({"a": pa.array([2.5], pa.float32()), "b": pa.array([7], pa.int64()), "s": pa.array(["x"])}),
    }
    for name, t in frames.items():
        for fn in ("global_summary", "measures_of_counts", "measures_of_centralTendency", "measures_of_cardinality",
                   "measures_of_dispersion", "measures_of_percentiles", "measures_of_shape", "missingCount_computation",
                   "nonzeroCount_computation", "mode_computation"):
            got, exp = getattr(sg, fn)(None, t).toPandas(), getattr(O, fn)(t)
            assert list(got.columns) == list(exp.columns) and len(got) == len(exp), (name, fn)
            for c in got.columns:
                for x, y in zip(got[c].tolist(), exp[c].tolist()):
                    assert (pd.isna(x) and (y is None or pd.isna(y))) or x == y or str(x) == str(y), (name, fn, c, x, y)
    # binning and drift on the all-null frame: column a is dropped with a warning, b is binned
    with pytest.warns(UserWarning):
        out = tr.attribute_binning(None, frames["all_null"], bin_size=3)
    assert out.columns == ["a", "b", "s"]
    r = dd.statistics(None, frames["all_null"], frames["all_null"], method_type="all", use_sampling=False,
                      source_path=str(tmp_path)).toPandas()
    assert r["attribute"].tolist() == ["a", "b", "s"] and r["flagged"].tolist() == [0, 0, 0]


def test_prefetch_pipeline_matches_direct(sg):
    """profile.prefetch (async grouped upload + passes) must give the same frames as the lazy path."""
    import torch
    from anovos_b200 import profile
    from anovos_b200.frame import ColumnFrame
    n = 300007
    rng = np.random.default_rng(5)
    host = {}
    for i in range(7):
        x = torch.from_numpy(rng.normal(i, 1 + i, n).astype(np.float32)).pin_memory()
        valid = rng.random(n) > 0.1 * (i % 3)
        bits = np.packbits(valid, bitorder="little")
        bits = np.concatenate([bits, np.zeros((-len(bits)) % 4, np.uint8)]).view(np.int32)
        host["c%d" % i] = (x, torch.from_numpy(bits).pin_memory()) if i % 3 else x
    a = ColumnFrame.from_tensors(host, n_rows=n)
    b = ColumnFrame.from_tensors(host, n_rows=n)
    profile.prefetch(b, group=3)
    for f in (sg.measures_of_counts, sg.measures_of_centralTendency, sg.measures_of_cardinality,
              sg.measures_of_dispersion, sg.measures_of_percentiles, sg.measures_of_shape):
        x, y = f(None, a).toPandas(), f(None, b).toPandas()
        assert x.equals(y), f.__name__


# ---- N1: stability_index_computation (drift_stability/test_stability.py:69-92, notebook cells 14-17) ---------

def test_stability_index(tmp_path, nb_drift):
    import anovos.drift_stability.stability as st
    from test_oracle_golden import _stab_tables, check_stability_notebook
    r = st.stability_index_computation(None, _stab_tables()).toPandas().iloc[0]
    np.testing.assert_almost_equal([r[c] for c in ("mean_cv", "stddev_cv", "kurtosis_cv", "mean_si", "stddev_si", "kurtosis_si",
                                                   "stability_index", "flagged")], [0.162, 0.62, 0.198, 2.0, 0.0, 2.0, 1.4, 0.0], 3)
    b = [pa.table({"A": np.array([0] * z + [1] * (20 - z))}) for z in (10, 12, 14)]
    r = st.stability_index_computation(None, b, binary_cols="A").toPandas().iloc[0]
    np.testing.assert_almost_equal([r["mean_stddev"], r["mean_si"], r["stability_index"], r["flagged"]], [0.1, 0.0, 0.0, 1.0], 3)
    with pytest.raises(ValueError):
        st.stability_index_computation(None, _stab_tables(), metric_weightages={"mean": 0.5})
    with pytest.raises(TypeError):
        st.stability_index_computation(None, _stab_tables(), binary_cols="Z")
    check_stability_notebook(lambda tables, **kw: st.stability_index_computation(None, tables, **kw).toPandas(), nb_drift, tmp_path)
    # product == oracle on every column of the 12-dataset run
    from test_oracle_golden import _stab_datasets
    ds = _stab_datasets()
    got = st.stability_index_computation(None, ds, threshold=2).toPandas()
    exp = O.stability_index_computation(ds, threshold=2)
    _cmp_frames(ResultLike(got), exp)


class ResultLike:
    def __init__(self, df):
        self.df = df

    def toPandas(self):
        return self.df


# ---- N2: quality_checker consumers (data_analyzer/test_quality_checker.py:252-524) ---------------------------

def test_quality_checker_consumers():
    import anovos.data_analyzer.quality_checker as qc
    t3 = O.table_from_rows([("27520a", 51, 9000, "HS-grad"), ("10a", 42, 7000, "Postgrad"), ("11a", 35, None, "graduate"),
                            ("1100b", 23, 6000, "matric")], ["ifa", "age", "income", "education"])
    odf, pr = qc.IDness_detection(None, t3, drop_cols=["ifa"], treatment=False, treatment_threshold=1.0)     # :278-307
    assert len(odf.columns) == 4
    e = _rec(pr, "education")
    assert e["unique_values"] == 4 and e["IDness"] == 1.0 and e["flagged"] == 1
    odf, pr = qc.IDness_detection(None, t3, drop_cols=["ifa"], treatment=True, treatment_threshold=1.0)      # :309-338
    assert len(odf.columns) == 1 and _rec(pr, "education")["treated"] == 1
    t4 = O.table_from_rows([("27520a", 51, 9000, "HS-grad"), ("10a", 42, 7000, "HS-grad"), ("11a", 35, None, "HS-grad"),
                            ("11d", 45, 9500, "HS-grad"), ("1100b", 23, 6000, "matric")], ["ifa", "age", "income", "education"])
    odf, pr = qc.biasedness_detection(None, t4, treatment=False, treatment_threshold=0.8)                    # :366-392
    e = _rec(pr, "education")
    assert len(odf.columns) == 4 and e["mode"] == "HS-grad" and e["mode_pct"] == 0.8 and e["flagged"] == 1
    odf, pr = qc.biasedness_detection(None, t4, treatment=True, treatment_threshold=0.8)                     # :394-420
    e = _rec(pr, "education")
    assert len(odf.columns) == 3 and e["mode"] == "HS-grad" and e["mode_pct"] == 0.8 and e["treated"] == 1
    t6 = O.table_from_rows([("27520a", 51, 9000, "HS-grad"), ("10a", 42, 7000, "Postgrad"), ("11a", 35, None, None),
                            ("1100b", 23, 6000, "HS-grad")], ["ifa", "age", "income", "education"])
    odf, pr = qc.nullColumns_detection(None, t6, treatment=True)                                             # :492-524
    assert len(odf.columns) == 4 and odf.count() == 3
    assert _rec(pr, "education")["missing_count"] == 1 and _rec(pr, "education")["missing_pct"] == 0.25
    assert _rec(pr, "income")["missing_count"] == 1 and _rec(pr, "income")["missing_pct"] == 0.25
    odf, pr = qc.nullColumns_detection(None, t6, treatment=True, treatment_method="column_removal",
                                       treatment_configs={"treatment_threshold": 0.2})
    assert odf.columns == ["ifa", "age"]
    with pytest.raises(TypeError):
        qc.nullColumns_detection(None, t6, treatment=True, treatment_method="column_removal")
    with pytest.raises(TypeError):
        qc.IDness_detection(None, t3, treatment="maybe")


def test_outlier_detection(income_part0, tmp_path):
    """data_analyzer/test_quality_checker.py:526-668 with unchanged inputs and expected values, plus product == oracle
    (thresholds, counts and treated frames) on a seeded frame with nulls."""
    import anovos.data_analyzer.quality_checker as qc
    from anovos_b200 import engine
    from anovos_b200.frame import ColumnFrame
    from test_oracle_golden import check_outlier_reference_tests

    def view(odf):
        def minmax(c):
            m = engine.moments(odf, [c])[0]
            return (m["min"], m["max"])
        return {"rows": odf.count(), "columns": odf.columns, "minmax": minmax,
                "nulls": lambda c: odf.count() - int(engine.moments(odf, [c])[0]["n_valid"])}

    def run(table, print_impact=False, **kw):
        r = qc.outlier_detection(None, table, print_impact=print_impact, **kw)
        return (view(r[0]), r[1].toPandas()) if print_impact else view(r)

    t = income_part0.append_column("label", pa.array([0] * income_part0.num_rows))
    check_outlier_reference_tests(run, t, tmp_path)
    # product vs oracle, every side / treatment, incl. a float32 column with NaN-free heavy tails and nulls
    rng = np.random.default_rng(11)
    n = 30_011
    x = rng.standard_t(3, n).astype(np.float32)
    y = rng.integers(0, 100, n).astype(np.int64)
    z = np.where(rng.random(n) < 0.97, 0.0, rng.exponential(3.0, n))
    tab = pa.table({"x": pa.array(x, mask=rng.random(n) < 0.1), "y": pa.array(y), "z": pa.array(z), "s": pa.array(["a"] * n)})
    for side in ("upper", "lower", "both"):
        for method in ("value_replacement", "null_replacement", "row_removal"):
            exp_t, exp_p = O.outlier_detection(tab, detection_side=side, treatment_method=method)
            got_t, got_p = qc.outlier_detection(None, tab, detection_side=side, treatment_method=method, print_impact=True)
            gp = got_p.toPandas()
            assert gp.values.tolist() == exp_p.values.tolist(), (side, method)
            assert got_t.count() == exp_t.num_rows and got_t.columns == exp_t.column_names
            for c in ("x", "y", "z"):
                d, v = got_t.column(c).device()
                e = exp_t.column(c).combine_chunks()
                valid = np.asarray(e.is_valid())
                g = d.cpu().numpy().astype(np.float64)
                assert np.array_equal(g[valid], np.asarray(e.fill_null(0)).astype(np.float64)[valid]), (side, method, c)
                gv = np.ones(len(g), bool) if v is None else ColumnFrame({c: got_t.column(c)}, got_t.count()).valid_mask(c).cpu().numpy()
                assert np.array_equal(gv, valid), (side, method, c)
    with pytest.raises(TypeError):
        qc.outlier_detection(None, tab, list_of_cols=["s"])
    with pytest.raises(TypeError):
        qc.outlier_detection(None, tab, detection_side="sideways")
    with pytest.raises(TypeError):
        qc.outlier_detection(None, tab, detection_configs={"pctile_upper": 1.5})
    with pytest.warns(UserWarning):
        assert qc.outlier_detection(None, tab, treatment=False, print_impact=False).count() == n


# ---- N3: IV / IG (data_analyzer/test_association_evaluator.py:25-240) ------------------------------------------

def test_iv_ig(income_part0):
    import anovos.data_analyzer.association_evaluator as ae
    from test_oracle_golden import check_iv_ig, label_table
    t = label_table(income_part0)
    iv, ig = ae.IV_calculation(None, t, drop_cols=["ifa"]).toPandas(), ae.IG_calculation(None, t, drop_cols=["ifa"]).toPandas()
    check_iv_ig(iv, ig)
    oiv, oig = O.IV_calculation(t, drop_cols=["ifa"]), O.IG_calculation(t, drop_cols=["ifa"])
    assert np.allclose(iv["iv"].values, oiv["iv"].values, rtol=1e-12, atol=1e-14)      # same counts -> same floats
    assert np.allclose(ig["ig"].values, oig["ig"].values, rtol=1e-12, atol=1e-14)
    er = ae.IV_calculation(None, t, list_of_cols=["age", "fnlwgt"], encoding_configs={"bin_method": "equal_range", "bin_size": 7,
                                                                                      "monotonicity_check": 0}).toPandas()
    eo = O.IV_calculation(t, list_of_cols=["age", "fnlwgt"], encoding_configs={"bin_method": "equal_range", "bin_size": 7,
                                                                               "monotonicity_check": 0})
    assert np.allclose(er["iv"].values, eo["iv"].values, rtol=1e-12)
    with pytest.raises(TypeError):
        ae.IV_calculation(None, t, label_col="nope")
    with pytest.raises(TypeError):
        ae.IG_calculation(None, t, event_label=7)
    # a string label works too
    ts = income_part0
    a = ae.IV_calculation(None, ts, list_of_cols=["sex", "age"], label_col="income", event_label=">50K").toPandas()
    assert abs(a.set_index("attribute").loc["sex", "iv"] - 0.3111) < 5e-5


def test_nb_iv_ig_with_spark_partitions(income_spark, nb_assoc):
    """The association notebook's stored IV / IG tables (108 values) through the product on the Spark-partitioned table:
    approxQuantile cutoffs from the merged per-partition sketches, label-class histograms from the binning kernels."""
    import functools
    import anovos.data_analyzer.association_evaluator as ae
    from test_oracle_golden import check_nb_iv_ig
    check_nb_iv_ig(lambda **kw: ae.IV_calculation(None, income_spark, **kw).toPandas(),
                   lambda **kw: ae.IG_calculation(None, income_spark, **kw).toPandas(), nb_assoc)


def test_nb_quality_checker(income, nb_quality):
    """Stored outputs of the quality-checker notebook on the income CSV: nullColumns (cells 17-19), IDness (35-37, HLL++
    default), biasedness (41-43; the mode VALUE only where it is unique - ties are arbitrary in the reference)."""
    import anovos.data_analyzer.quality_checker as qc
    calls = {
        17: (qc.nullColumns_detection, {}), 18: (qc.nullColumns_detection, {"list_of_cols": "all", "drop_cols": ["ifa"]}),
        19: (qc.nullColumns_detection, {"list_of_cols": ["age", "sex", "race", "workclass", "fnlwgt"]}),
        35: (qc.IDness_detection, {}), 36: (qc.IDness_detection, {"list_of_cols": "all", "drop_cols": ["ifa"], "treatment_threshold": 0.75}),
        37: (qc.IDness_detection, {"list_of_cols": ["sex", "race", "workclass"]}),
        41: (qc.biasedness_detection, {}),
        42: (qc.biasedness_detection, {"list_of_cols": "all", "drop_cols": ["ifa"], "treatment_threshold": 0.75}),
        43: (qc.biasedness_detection, {"list_of_cols": ["age", "sex", "race", "workclass", "logfnl"]}),
    }
    checked = 0
    for cell, (fn, kw) in calls.items():
        _, pr = fn(None, income, **kw)
        got, exp = frame_by_attr(pr.toPandas()), table_by_attr(nb_quality[cell])
        assert set(got) == set(exp), (cell, sorted(set(got) ^ set(exp)))
        for a, row in exp.items():
            for c, shown in row.items():
                if c in ("attribute", "mode"):
                    continue
                g = got[a][c]
                assert shown_close(None if pd.isna(g) else g, shown), (cell, a, c, g, shown)
                checked += 1
    assert checked > 250
    _, pr = qc.biasedness_detection(None, income)
    got, exp = frame_by_attr(pr.toPandas()), table_by_attr(nb_quality[41])
    same_mode = sum((str(got[a]["mode"]) == exp[a]["mode"]) or (pd.isna(got[a]["mode"]) and exp[a]["mode"] in ("None", "NaN"))
                    for a in exp)
    assert same_mode >= len(exp) - 3          # ties (fnlwgt: 13 rows on several values; ifa: every id once) are arbitrary


def test_nb_attribute_binning_head(tr, income_spark):
    """transformers notebook cells 6 / 8 through the product (the binned frame of a partitioned table is partitioned)."""
    from test_oracle_golden import NB_BINNING
    for method, exp in NB_BINNING.items():
        out = tr.attribute_binning(None, income_spark, list_of_cols=["education-num", "hours-per-week"], method_type=method, bin_size=5)
        first = next(iter(out.chunks(["education-num", "hours-per-week"])))
        got = [[int(first.column(c).device()[0][i].item()) for c in ("education-num", "hours-per-week")] for i in range(5)]
        assert got == exp, (method, got)



def test_percentiles_of_partitions_beyond_the_head_buffer(sg):
    """Spark partitions of >= 50 000 non-null values: the device sorts every 50 000-value head-buffer batch, the library's
    host helper runs Spark's sequential merge / compress -> product == the oracle's sketch on every percentile, median,
    IQR; float64 with nulls, float32, int32 and a mostly-null column that stays under the head size."""
    from test_host_paths_cpu import _large_partition_table
    t, parts = _large_partition_table()
    tt = O.with_spark_partitions(t, parts)
    got = sg.measures_of_percentiles(None, tt).toPandas()
    exp = O.measures_of_percentiles(tt)
    assert got.equals(exp), (got, exp)
    assert sg.measures_of_dispersion(None, tt).toPandas()["IQR"].tolist() == O.measures_of_dispersion(tt)["IQR"].tolist()
    assert sg.measures_of_centralTendency(None, tt).toPandas()["median"].tolist() == O.measures_of_centralTendency(tt)["median"].tolist()
    assert not O.measures_of_percentiles(t).equals(exp)


def test_narrow_host_codes_upload_like_int32(sg):
    """String columns keep their dictionary codes on the host in uint8 / int16 / int32 by cardinality; lazy upload and the
    prefetch pipeline widen them on the device, and every statistic equals the one of the same frame with int32 host codes
    and the oracle's."""
    import torch
    from anovos_b200 import profile
    from anovos_b200.frame import ColumnFrame, narrow_code_dtype
    n = 200_003
    rng = np.random.default_rng(9)
    cards = [2, 200, 256, 257, 30_000, 40_000]
    cols = {"s%d" % k: pa.array(["v%05d" % v for v in rng.integers(0, k, n)], mask=rng.random(n) < 0.05) for k in cards}
    cols["x"] = pa.array(rng.normal(0, 1, n).astype(np.float32))
    t = pa.table(cols)
    fr = ColumnFrame.from_arrow(t)
    for k in cards:
        c = fr.column("s%d" % k)
        assert c._host.dtype == narrow_code_dtype(len(c.dictionary))
    assert [fr.column("s%d" % k)._host.dtype.itemsize for k in cards] == [1, 1, 1, 2, 2, 4]
    wide = {}
    for name in fr.columns:
        c = fr.column(name)
        h = torch.from_numpy(c._host.astype(np.int32) if c.dictionary is not None else c._host)
        hv = None if c._host_valid is None else torch.from_numpy(c._host_valid)
        wide[name] = (h, hv, c.dictionary) if c.dictionary is not None else ((h, hv) if hv is not None else h)
    narrow = {name: ((torch.from_numpy(fr.column(name)._host).pin_memory(),) + tuple(v[1:]) if isinstance(v, tuple) and len(v) == 3 else v)
              for name, v in wide.items()}
    a = ColumnFrame.from_tensors(wide, n_rows=n)
    b = ColumnFrame.from_tensors(narrow, n_rows=n)
    profile.prefetch(b, group=3)
    for name in fr.columns:
        d, _ = fr.column(name).device()
        d2, _ = b.column(name).device()
        if fr.column(name).dictionary is not None:
            assert d.dtype == torch.int32 and d2.dtype == torch.int32
            assert torch.equal(d, a.column(name).device()[0]) and torch.equal(d2, d)
    for f in ("measures_of_counts", "measures_of_centralTendency", "measures_of_cardinality"):
        x, y, z = (getattr(sg, f)(None, q).toPandas() for q in (a, b, fr))
        assert x.equals(y) and x.equals(z), f
        exp = getattr(O, f)(t)
        from golden_util import assert_frames_match
        assert_frames_match(x, exp)
    assert fr.to_arrow().equals(t)
