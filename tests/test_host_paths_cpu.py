"""Host paths of the product under `-m "not gpu"`: the NumPy engine stand-in of tests/cpu_engine.py replaces the kernels, so
drift (key alignment, saved model / frequency files, flags), outlier thresholds, Spark-partitioned percentiles and the
row-partition merges run end to end on CPU and are compared with the oracle and with the reference's pins."""
import json
import os
import warnings

import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

import cpu_engine
from golden_util import frame_by_attr, shown_close, table_by_attr
from oracle import api as O
from oracle import spark_semantics as S


def _close_cols(a, b, cols, rtol=1e-9):
    for c in cols:
        assert np.allclose(np.asarray(a[c], float), np.asarray(b[c], float), rtol=rtol, atol=1e-15, equal_nan=True), c


def _drift_tables():
    rng = np.random.default_rng(4)
    n = 12_007
    def mk(shift, seed):
        r = np.random.default_rng(seed)
        return pa.table({
            "x": pa.array(r.normal(shift, 2, n).astype(np.float32), mask=r.random(n) < 0.03),
            "k": pa.array(r.integers(0, 40, n).astype(np.int64)),
            "all_null": pa.array([None] * n, pa.float64()),
            "s": pa.array(r.choice(["a", "b", "c", "dd", "e,f"], n, p=[0.4, 0.3, 0.2, 0.05, 0.05]), mask=r.random(n) < 0.1),
            "wide": pa.array(["k%04d" % v for v in r.integers(0, 300, n)]),
        })
    return mk(0.0, 1), mk(0.4, 2)


@pytest.mark.parametrize("bin_method", ["equal_range", "equal_frequency"])
def test_drift_host_path_equals_oracle(bin_method, tmp_path):
    import anovos.drift_stability.drift_detector as dd
    src, tgt = _drift_tables()
    kw = dict(method_type="all", use_sampling=False, bin_method=bin_method, bin_size=8)
    with cpu_engine.installed(), warnings.catch_warnings():
        warnings.simplefilter("ignore")          # equal_range drops the all-null column with a warning
        got = dd.statistics(None, tgt, src, source_path=str(tmp_path / "p"), **kw).toPandas()
        again = dd.statistics(None, tgt, None, pre_existing_source=True, source_path=str(tmp_path / "p"), **kw).toPandas()
        exp = O.statistics(tgt, src, source_path=str(tmp_path / "o"), **kw)
    assert got["attribute"].tolist() == exp["attribute"].tolist() and got["flagged"].tolist() == exp["flagged"].tolist()
    _close_cols(got, exp, ["PSI", "HD", "JSD", "KS"])
    _close_cols(again, exp, ["PSI", "HD", "JSD", "KS"], rtol=1e-9)     # the saved model + frequency CSVs round-trip
    # artefacts in the reference's layout
    assert os.path.isdir(tmp_path / "p" / "drift_statistics" / "attribute_binning")
    f = pd.read_csv(tmp_path / "p" / "drift_statistics" / "frequency_counts" / "s" / "part-00000.csv")
    assert list(f.columns) == ["s", "p"] and abs(f["p"].sum() - (1 - src.column("s").null_count / src.num_rows)) < 1e-12


def test_nb_drift_and_percentiles_through_the_host_path(income, income_source, income_spark, nb_drift, nb_stats, tmp_path):
    """Notebook pins through the PRODUCT's host code on CPU: the 21-column PSI table and - with Spark's partitioning -
    all 81 summary() percentiles (per-partition sketch samples merged by shared/gk.py)."""
    import anovos.data_analyzer.stats_generator as sg
    import anovos.drift_stability.drift_detector as dd
    with cpu_engine.installed():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d = dd.statistics(None, income, income_source, method_type="PSI", use_sampling=False, source_path=str(tmp_path)).toPandas()
        pct = sg.measures_of_percentiles(None, income_spark).toPandas()
        cen = sg.measures_of_centralTendency(None, income_spark).toPandas()
    psi_tables = [t for t in nb_drift.values() if t["columns"][:2] == ["attribute", "PSI"] and len(t["rows"]) >= 20]
    exp = table_by_attr(psi_tables[0])
    got = frame_by_attr(d)
    assert sum(shown_close(got[a]["PSI"], r["PSI"]) for a, r in exp.items() if a in got) >= len(exp) - 1
    exp = table_by_attr(nb_stats[35])
    got = frame_by_attr(pct)
    for a, row in exp.items():
        for c in ["min", "max"] + list(S.SUMMARY_PCTS):
            g = got[a][c]
            assert shown_close(None if pd.isna(g) else g, row[c]), (a, c, g, row[c])
    exp, got = table_by_attr(nb_stats[17]), frame_by_attr(cen)
    for a, row in exp.items():
        g = got[a]["median"]
        assert shown_close(None if pd.isna(g) else g, row["median"]), (a, g, row["median"])


def test_outlier_thresholds_and_counts_host_path(income_part0):
    """The 13 reference pins of test_quality_checker.py:526-637 that do not need a treated frame: thresholds from
    approxQuantile ranks / moments, counts from the compare pass, skew exclusion, saved-model round trip."""
    import anovos.data_analyzer.quality_checker as qc
    from test_oracle_golden import OUTLIER_PINS_UPPER
    t = income_part0.append_column("label", pa.array([0] * income_part0.num_rows))
    with cpu_engine.installed(), warnings.catch_warnings():
        warnings.simplefilter("ignore")          # capital-loss is excluded as skewed, with a warning
        _, pr = qc.outlier_detection(None, t, drop_cols=["ifa", "label"], treatment=False, print_impact=True)
    got = {r["attribute"]: [r["lower_outliers"], r["upper_outliers"], r["excluded_due_to_skewness"]] for r in pr.toPandas().to_dict("records")}
    assert got == OUTLIER_PINS_UPPER
    with cpu_engine.installed():
        _, pr = qc.outlier_detection(None, t, list_of_cols=["age", "education-num"], detection_side="both",
                                     detection_configs={"pctile_lower": 0.02, "pctile_upper": 0.98}, treatment=False, print_impact=True)
    got = {r["attribute"]: [r["lower_outliers"], r["upper_outliers"]] for r in pr.toPandas().to_dict("records")}
    assert got == {"age": [202, 482], "education-num": [267, 205]}


def test_row_partition_merges_host_path(income):
    """PartitionedFrame on CPU: chunked moments / histograms / code counts / HLL registers / percentiles == the whole frame."""
    import anovos.data_analyzer.stats_generator as sg
    from anovos_b200.frame import ColumnFrame
    from anovos_b200.partitioned import PartitionedFrame
    with cpu_engine.installed():
        whole = ColumnFrame.from_arrow(income)
        parts = PartitionedFrame.from_frame(income, 4096)
        assert parts.n_chunks == 8 and parts.count() == income.num_rows
        for fn in ("measures_of_counts", "measures_of_percentiles", "measures_of_cardinality", "measures_of_centralTendency"):
            a, b = getattr(sg, fn)(None, whole).toPandas(), getattr(sg, fn)(None, parts).toPandas()
            assert a.equals(b), fn
        for fn in ("measures_of_dispersion", "measures_of_shape"):
            a, b = getattr(sg, fn)(None, whole).toPandas(), getattr(sg, fn)(None, parts).toPandas()
            assert np.allclose(a.drop(columns="attribute").to_numpy(float), b.drop(columns="attribute").to_numpy(float),
                               rtol=1e-9, atol=1.01e-4, equal_nan=True), fn


def test_stability_index_host_path(nb_drift, tmp_path):
    """stability_index_computation (test_stability.py:69-92 and notebook cells 14-17): K1 per dataset comes from the
    stand-in, the CV / score / weighting arithmetic and the metric files are the product's."""
    import anovos.drift_stability.stability as st
    from test_oracle_golden import _stab_tables, check_stability_notebook
    with cpu_engine.installed():
        r = st.stability_index_computation(None, _stab_tables()).toPandas().iloc[0]
        np.testing.assert_almost_equal([r[c] for c in ("mean_cv", "stddev_cv", "kurtosis_cv", "mean_si", "stddev_si", "kurtosis_si",
                                                       "stability_index", "flagged")], [0.162, 0.62, 0.198, 2.0, 0.0, 2.0, 1.4, 0.0], 3)
        with pytest.raises(ValueError):
            st.stability_index_computation(None, _stab_tables(), metric_weightages={"mean": 0.5})
        check_stability_notebook(lambda tables, **kw: st.stability_index_computation(None, tables, **kw).toPandas(), nb_drift, tmp_path)


def test_quality_checker_notebook_host_path(income, nb_quality):
    """nullColumns / IDness (HLL++ default) / biasedness tables of the quality-checker notebook through the product's
    host code (cells 17-19, 35-37, 41-43)."""
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
    with cpu_engine.installed():
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


def test_attribute_binning_host_path(income_part1, tmp_path):
    """data_transformer/test_transformers.py:37-104 with the reference's inputs and checks, plus ids == oracle for both
    methods, the saved-model round trip, categorical labels and the append mode."""
    import anovos.data_transformer.transformers as tr
    cols = ["age", "fnlwgt", "hours-per-week"]

    def ids(fr, c):
        d, v = fr.column(c).device()
        return np.asarray(d), v

    with cpu_engine.installed():
        out = tr.attribute_binning(None, income_part1, list_of_cols=cols, bin_size=20, model_path=str(tmp_path))
        assert len(out.columns) == 17
        for c in cols:
            d, v = ids(out, c)
            valid = np.asarray(income_part1.column(c).is_valid())
            assert d[valid].min() == 1 and d[valid].max() == 20
        with pytest.raises(IndexError, match="list index out of range"):
            tr.attribute_binning(None, income_part1, list_of_cols=["education-num"], bin_size=20, pre_existing_model=True,
                                 model_path=str(tmp_path))
        same = tr.attribute_binning(None, income_part1, list_of_cols=[], bin_size=20)
        assert same.columns == income_part1.column_names
        app = tr.attribute_binning(None, income_part1, list_of_cols=cols, bin_size=20, output_mode="append")
        assert len(app.columns) == 20 and app.columns[-3:] == [c + "_binned" for c in cols]
        for method in ("equal_range", "equal_frequency"):
            got = tr.attribute_binning(None, income_part1, list_of_cols=cols, method_type=method, bin_size=7)
            exp = O.attribute_binning(income_part1, list_of_cols=cols, method_type=method, bin_size=7)
            for c in cols:
                d, _ = ids(got, c)
                e = np.asarray(exp.column(c).fill_null(0))
                assert np.array_equal(d, e), (method, c)
        again = tr.attribute_binning(None, income_part1, list_of_cols=cols, pre_existing_model=True, model_path=str(tmp_path))
        for c in cols:
            assert np.array_equal(ids(again, c)[0], ids(out, c)[0])
        lab = tr.attribute_binning(None, income_part1, list_of_cols=["age"], bin_size=4, bin_dtype="categorical")
        assert lab.column("age").kind == "cat" and len(lab.column("age").dictionary) == 4
        assert lab.column("age").dictionary[0].startswith("<= ") and lab.column("age").dictionary[-1].startswith("> ")
        for bad in (dict(bin_size=1), dict(method_type="foo"), dict(output_mode="x"), dict(list_of_cols=["workclass"])):
            with pytest.raises(TypeError):
                tr.attribute_binning(None, income_part1, **{"list_of_cols": cols, **bad})


def test_outlier_treatments_host_path(income_part0, tmp_path):
    """All four reference outlier tests (test_quality_checker.py:526-668) incl. the treated frames - row removal, value
    replacement (clamped min / max), null replacement with a saved model - through the product's host code; the per-row
    compare comes from the stand-in's bin ids, the treated columns from the product's tensor ops (CPU tensors here)."""
    import anovos.data_analyzer.quality_checker as qc
    from anovos_b200 import engine
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
    with cpu_engine.installed(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        check_outlier_reference_tests(run, t, tmp_path)


def test_iv_ig_host_path(income_part0, income_spark, nb_assoc):
    """IV / IG through the product's host code: the unit-test pins of test_association_evaluator.py:49-240 (single parquet
    file) and the 108 stored notebook values (income CSV with Spark's two scan partitions)."""
    import functools
    import anovos.data_analyzer.association_evaluator as ae
    from test_oracle_golden import check_iv_ig, check_nb_iv_ig, label_table
    with cpu_engine.installed():
        t = label_table(income_part0)
        check_iv_ig(ae.IV_calculation(None, t, drop_cols=["ifa"]).toPandas(), ae.IG_calculation(None, t, drop_cols=["ifa"]).toPandas())
        check_nb_iv_ig(lambda **kw: ae.IV_calculation(None, income_spark, **kw).toPandas(),
                       lambda **kw: ae.IG_calculation(None, income_spark, **kw).toPandas(), nb_assoc)
        with pytest.raises(TypeError):
            ae.IV_calculation(None, t, label_col="nope")
        with pytest.raises(TypeError):
            ae.IV_calculation(None, t, event_label=7)


def test_saved_frequency_round_trip_keeps_string_keys(tmp_path):
    """pre_existing_source=True must join on the strings that were saved: categories spelled like pandas' NA markers
    ("NA", "None", "null", "nan", "N/A") or like numbers ("007", "1.0") are ordinary categories for Spark's CSV reader."""
    import anovos.drift_stability.drift_detector as dd
    rng = np.random.default_rng(3)
    n = 6000
    cats = ["NA", "None", "null", "nan", "N/A", "007", "1.0", "plain", "a,b"]
    def mk(p, seed):
        r = np.random.default_rng(seed)
        return pa.table({"s": pa.array(r.choice(cats, n, p=p), mask=r.random(n) < 0.05), "x": pa.array(r.normal(0, 1, n))})
    p0 = np.array([3, 2, 2, 1, 1, 2, 2, 4, 1], float); p0 /= p0.sum()
    p1 = np.array([1, 3, 1, 2, 1, 3, 1, 3, 2], float); p1 /= p1.sum()
    src, tgt = mk(p0, 1), mk(p1, 2)
    kw = dict(method_type="all", use_sampling=False)
    with cpu_engine.installed():
        direct = dd.statistics(None, tgt, src, source_path=str(tmp_path / "m"), **kw).toPandas()
        again = dd.statistics(None, tgt, None, pre_existing_source=True, source_path=str(tmp_path / "m"), **kw).toPandas()
    exp = O.statistics(tgt, src, source_path=str(tmp_path / "o"), **kw)
    _close_cols(direct, exp, ["PSI", "HD", "JSD", "KS"])
    _close_cols(again, direct, ["PSI", "HD", "JSD", "KS"])
    f = pd.read_csv(tmp_path / "m" / "drift_statistics" / "frequency_counts" / "s" / "part-00000.csv", dtype=str, keep_default_na=False)
    assert sorted(k for k in f["s"] if k != "") == sorted(cats)


def test_outlier_nan_values_are_not_outliers(tmp_path):
    """`(v - upper) > 0` is False for NaN in the reference's compare (quality_checker.py:937-966): a NaN value is never
    flagged, although the binning kernels put it in the last bin.  Thresholds come from a saved model (NaNs in the data the
    thresholds are computed from poison them in Spark as well: unpinned)."""
    from anovos_b200.data_analyzer import quality_checker as qc
    rng = np.random.default_rng(2)
    x = rng.normal(0, 1, 4000)
    clean = pa.table({"x": pa.array(x), "y": pa.array(rng.normal(0, 1, 4000))})
    x2 = x.copy()
    x2[:7] = np.nan
    x2[7:12] = 50.0
    dirty = pa.table({"x": pa.array(x2), "y": clean.column("y")})
    with cpu_engine.installed(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        qc.outlier_detection(None, clean, detection_side="both", treatment=False, print_impact=True, model_path=str(tmp_path))
        base = qc.outlier_detection(None, clean, detection_side="both", pre_existing_model=True, model_path=str(tmp_path),
                                    print_impact=True, treatment=False)[1].toPandas().set_index("attribute")
        odf, imp = qc.outlier_detection(None, dirty, detection_side="both", pre_existing_model=True, model_path=str(tmp_path),
                                        print_impact=True, treatment=True, treatment_method="null_replacement")
        imp_only = qc.outlier_detection(None, dirty, detection_side="both", pre_existing_model=True, model_path=str(tmp_path),
                                        print_impact=True, treatment=False)[1].toPandas().set_index("attribute")
        v = odf.column("x").device()[1]
    r = imp.toPandas().set_index("attribute")
    clean_upper_in_first12 = int((x[:12] > 0).sum())      # at most the ordinary tail values that were overwritten
    assert r.loc["x", "upper_outliers"] == imp_only.loc["x", "upper_outliers"]                    # rows path == histogram path
    assert base.loc["x", "upper_outliers"] + 5 - clean_upper_in_first12 <= r.loc["x", "upper_outliers"] <= base.loc["x", "upper_outliers"] + 5
    valid = np.unpackbits(v.numpy().view(np.uint8), bitorder="little")[:4000].astype(bool)
    assert valid[:7].all() and not valid[7:12].any()             # NaNs stay, the five 50.0s became null


def _large_partition_table():
    rng = np.random.default_rng(3)
    n = 260_000
    t = pa.table({"x": pa.array(np.round(rng.normal(10, 50, n), 1), mask=rng.random(n) < 0.05),
                  "f": pa.array(rng.lognormal(0, 1, n).astype(np.float32)),
                  "i": pa.array(rng.integers(0, 1000, n).astype(np.int32)),
                  "mostly_null": pa.array(rng.normal(0, 1, n), mask=rng.random(n) < 0.9)})
    return t, [120_001, 49_999, 60_000, 30_000]


@pytest.mark.parametrize("n", [0, 1, 2, 777, 49_999, 50_000, 50_001, 120_001, 260_007])
@pytest.mark.parametrize("eps", [1e-4, 0.01])
def test_gk_partition_sketch_helper_equals_oracle(n, eps):
    """anv_gk_partition_sketch (host helper of the C ABI: Spark's head-buffer flush + compress over batches the device has
    sorted) returns the oracle's sketch of the same arrival-ordered values, sample for sample."""
    import ctypes as C
    from anovos_b200 import _lib
    L = _lib.lib()
    v = np.round(np.random.default_rng(n).normal(0, 100, n), 1)
    H = 50_000
    sb = np.concatenate([np.sort(v[i:i + H]) for i in range(0, n, H)]) if n else np.zeros(0)
    cap = n + 8
    ov, og, od = np.zeros(cap), np.zeros(cap, np.int64), np.zeros(cap, np.int64)
    k = L.anv_gk_partition_sketch(sb.ctypes.data_as(C.c_void_p), n, H, eps, 10_000, ov.ctypes.data_as(C.c_void_p),
                                  og.ctypes.data_as(C.c_void_p), od.ctypes.data_as(C.c_void_p), cap)
    exp, cnt = S.gk_sketch(v, eps)
    assert cnt == n and k == len(exp)
    assert [(a, int(b), int(c)) for a, b, c in zip(ov[:k], og[:k], od[:k])] == [(float(a), int(b), int(c)) for a, b, c in exp]
    if n > 4:      # capacity too small: error code, nothing written past the end
        assert L.anv_gk_partition_sketch(sb.ctypes.data_as(C.c_void_p), n, H, eps, 10_000, ov.ctypes.data_as(C.c_void_p),
                                         og.ctypes.data_as(C.c_void_p), od.ctypes.data_as(C.c_void_p), 2) < 0


def test_percentiles_of_partitions_beyond_the_head_buffer_host_path():
    """Partitions of >= 50 000 non-null values (Spark flushes its head buffer as rows arrive): product == the oracle's
    sketch for every percentile, the median and the IQR; and the result is NOT the exact order statistic everywhere."""
    import anovos.data_analyzer.stats_generator as sg
    t, parts = _large_partition_table()
    tt = O.with_spark_partitions(t, parts)
    with cpu_engine.installed():
        got = sg.measures_of_percentiles(None, tt).toPandas()
        disp = sg.measures_of_dispersion(None, tt).toPandas()
        cen = sg.measures_of_centralTendency(None, tt).toPandas()
    exp = O.measures_of_percentiles(tt)
    assert got.equals(exp)
    assert disp["IQR"].tolist() == O.measures_of_dispersion(tt)["IQR"].tolist()
    assert cen["median"].tolist() == O.measures_of_centralTendency(tt)["median"].tolist()
    assert not O.measures_of_percentiles(t).equals(exp)
