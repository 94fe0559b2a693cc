"""The oracle against every golden vector the reference holds for the hot path:
the reference's own unit-test constants (cited file:line, relative to
/root/reference/src/test/anovos) and the stored Spark outputs of its notebooks."""
import math
import tempfile

import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

from oracle import api as O
from oracle import spark_semantics as S
from golden_util import shown_close, frame_by_attr, table_by_attr, cell_value


def _df4():
    return O.table_from_rows([("27520a", 51, "HS-grad"), ("10a", 42, "Postgrad"), ("11a", 55, None),
                              ("1100b", 23, "HS-grad")], ["ifa", "age", "education"])


# ---- data_analyzer/test_stats_generator.py ---------------------------------

def test_missing_count():  # :29-65
    r = frame_by_attr(O.missingCount_computation(_df4()))
    assert len(r) == 3
    assert r["education"]["missing_count"] == 1 and r["education"]["missing_pct"] == 0.25


def test_unique_count():  # :68-184
    t = O.table_from_rows([("27520a", 51, 9000, "HS-grad"), ("10a", 42, 7000, "Postgrad"),
                           ("11a", 35, None, None), ("1100b", 23, 6000, "HS-grad")],
                          ["ifa", "age", "income", "education"])
    for kw in ({}, {"compute_approx_unique_count": True}, {"compute_approx_unique_count": True, "rsd": 0.05},
               {"compute_approx_unique_count": True, "rsd": 0.2}):
        r = frame_by_attr(O.uniqueCount_computation(t, **kw))
        assert r["education"]["unique_values"] == 2 and r["age"]["unique_values"] == 4
        assert r["income"]["unique_values"] == 3
    with pytest.raises(ValueError):
        O.uniqueCount_computation(t, compute_approx_unique_count=True, rsd=-0.1)


def test_mode():  # :187-235
    t = O.table_from_rows([("27520a", 51, "HS-grad"), ("10a", 42, "Postgrad"), ("11a", 55, None),
                           ("13a", 42, "HS-grad"), ("1100b", 23, "HS-grad")], ["ifa", "age", "education"])
    r = frame_by_attr(O.mode_computation(t))
    assert len(r) == 3
    assert r["education"]["mode"] == "HS-grad" and r["education"]["mode_rows"] == 3
    assert r["age"]["mode"] == "42" and r["age"]["mode_rows"] == 2


def test_nonzero():  # :238-289
    t = O.table_from_rows([("27520a", 51, 9000, "HS-grad"), ("10a", 0, 7000, "Postgrad"),
                           ("11a", 35, None, None), ("1100b", 23, 6000, "HS-grad")],
                          ["ifa", "age", "income", "education"])
    r = frame_by_attr(O.nonzeroCount_computation(t))
    assert r["age"]["nonzero_count"] == 3 and r["age"]["nonzero_pct"] == 0.75
    assert r["income"]["nonzero_count"] == 3 and r["income"]["nonzero_pct"] == 0.75


def test_central_tendency():  # :292-339
    t = O.table_from_rows([("27520a", 51, "HS-grad"), ("10a", 42, "Postgrad"), ("11a", 55, None),
                           ("1100b", 23, "HS-grad"), ("1100c", 23, "HS-grad")][:4], ["ifa", "age", "education"])
    r = frame_by_attr(O.measures_of_centralTendency(t))
    assert r["age"]["mean"] == 42.75 and r["age"]["median"] == 42.0
    assert r["education"]["mode"] == "HS-grad" and r["education"]["mode_pct"] == 0.6667


def test_cardinality():  # :342-448
    r = frame_by_attr(O.measures_of_cardinality(_df4()))
    assert r["age"]["IDness"] == 1.0 and r["education"]["IDness"] == 0.6667
    assert r["education"]["unique_values"] == 2
    r = frame_by_attr(O.measures_of_cardinality(_df4(), use_approx_unique_count=False))
    assert r["age"]["IDness"] == 1.0 and r["education"]["IDness"] == 0.6667


def test_dispersion():  # :451-504
    r = frame_by_attr(O.measures_of_dispersion(_df4()))["age"]
    assert (r["stddev"], r["variance"], r["cov"], r["IQR"], r["range"]) == (14.2449, 202.9172, 0.3332, 28.0, 32.0)


def test_counts():  # :508-567
    r = frame_by_attr(O.measures_of_counts(_df4()))["age"]
    assert (r["fill_count"], r["fill_pct"], r["missing_count"], r["missing_pct"],
            r["nonzero_count"], r["nonzero_pct"]) == (4, 1.0, 0, 0.0, 4, 1.0)


def test_shape():  # :570-605
    r = frame_by_attr(O.measures_of_shape(_df4()))["age"]
    assert r["skewness"] == -0.7063 and r["kurtosis"] == -1.0646


def test_global_summary():  # :608-661
    g = dict(O.global_summary(_df4()).values.tolist())
    assert g["rows_count"] == "4" and g["columns_count"] == "3" and g["numcols_count"] == "1"
    assert g["numcols_name"] == "age" and g["catcols_count"] == "2"


def test_percentiles_small():  # :664-779 (inequalities in the reference)
    t = O.table_from_rows([("a", 51), ("b", 42), ("c", 55), ("d", 23), ("e", 46), ("f", 33)], ["ifa", "age"])
    r = frame_by_attr(O.measures_of_percentiles(t))["age"]
    assert r["min"] == 23.0 and r["max"] == 55.0 and r["50%"] <= 46.0 and r["1%"] <= 25.0


# ---- drift_stability/test_drift_detector.py:7-46 ----------------------------

def test_drift_known_answer(tmp_path):
    r = np.array([0.34, -1.76, 0.32, -0.39, -0.67, 0.61, 1.03, 0.93, -0.84, -0.31])
    tgt = pa.table({"A": r, "B": r})
    src = pa.table({"A": r, "B": r + 1})
    d = frame_by_attr(O.statistics(tgt, src, method_type="all", source_path=str(tmp_path)))
    assert [d["A"][k] for k in ("PSI", "HD", "JSD", "KS")] == [0, 0, 0, 0]
    assert [d["A"]["flagged"], d["B"]["flagged"]] == [0, 1]
    np.testing.assert_almost_equal([d["B"][k] for k in ("PSI", "HD", "JSD", "KS")],
                                   [7.6776, 0.7091, 0.3704, 0.4999], 4)
    assert list(O.statistics(tgt, src, method_type="all", source_path=str(tmp_path)).columns) == \
        ["attribute", "PSI", "HD", "JSD", "KS", "flagged"]
    e = frame_by_attr(O.statistics(tgt, src, method_type="all", bin_method="equal_frequency",
                                   source_path=str(tmp_path)))
    assert [e["A"][k] for k in ("PSI", "HD", "JSD", "KS")] == [0, 0, 0, 0]
    np.testing.assert_almost_equal([e["B"][k] for k in ("PSI", "HD", "JSD", "KS")],
                                   [3.0899, 0.4775, 0.1769, 0.4], 4)
    assert [e["A"]["flagged"], e["B"]["flagged"]] == [0, 1]


def test_drift_validation_errors(tmp_path):  # test_validations.py:13-20
    r = np.arange(5.0)
    t = pa.table({"A": r})
    with pytest.raises(ValueError):
        O.statistics(t, t, list_of_cols=[], source_path=str(tmp_path))
    with pytest.raises(ValueError):
        O.statistics(t, t, list_of_cols=["A"], drop_cols=["A"], source_path=str(tmp_path))
    with pytest.raises(TypeError):
        O.statistics(t, t, method_type="XYZ", source_path=str(tmp_path))


# ---- data_transformer/test_transformers.py:39-104 ---------------------------

def test_attribute_binning_income(income_part1, tmp_path):
    cols = ["age", "fnlwgt", "hours-per-week"]
    out = O.attribute_binning(income_part1, list_of_cols=cols, bin_size=20, model_path=str(tmp_path))
    for c in cols:
        v = np.asarray(out.column(c).drop_null())
        assert v.min() == 1 and v.max() == 20
    assert out.column("education-num").equals(income_part1.column("education-num"))
    app = O.attribute_binning(income_part1, list_of_cols=cols, bin_size=20, output_mode="append")
    assert len(app.column_names) == len(income_part1.column_names) + 3
    with pytest.raises(IndexError):
        O.attribute_binning(income_part1, list_of_cols=["capital-gain"], bin_size=20, pre_existing_model=True,
                            model_path=str(tmp_path))
    with pytest.raises(TypeError):
        O.attribute_binning(income_part1, list_of_cols=cols, bin_size=1)
    with pytest.raises(TypeError):
        O.attribute_binning(income_part1, list_of_cols=cols, method_type="foo")


# ---- notebook golden vectors (real Spark outputs on the income CSV) ---------

def _check_table(df, t, cols, skip=()):
    got, exp = frame_by_attr(df), table_by_attr(t)
    assert set(got) == set(exp)
    bad = []
    for a, row in exp.items():
        for c in cols:
            if (a, c) in skip:
                continue
            if not shown_close(got[a][c], row[c]):
                bad.append((a, c, got[a][c], row[c]))
    assert not bad, bad


def _rank_close(p, v, q, eps=1e-4):
    """v (shown to 4 decimals) must be a data element within eps*n ranks of ceil(q*n)."""
    srt = np.array([S.round_half_up(x) for x in p.sorted64])
    lo = np.searchsorted(srt, v - 5.1e-5, "left") + 1
    hi = np.searchsorted(srt, v + 5.1e-5, "right")
    assert hi >= lo, (p.name, q, v)
    r = S.quantile_rank(q, p.n)
    dist = 0 if lo <= r <= hi else min(abs(r - lo), abs(r - hi))
    assert dist <= eps * p.n + 1, (p.name, q, v, r, lo, hi)
    return dist


def test_nb_counts(income, nb_stats):
    _check_table(O.measures_of_counts(income), nb_stats[11],
                 ["fill_count", "fill_pct", "missing_count", "missing_pct", "nonzero_count", "nonzero_pct"])


def test_nb_central_tendency(income, nb_stats):
    df = O.measures_of_centralTendency(income)
    # mode ties are arbitrary in the reference (fnlwgt: 13 rows on two values) -> compare rows, not value
    _check_table(df, nb_stats[17], ["mean", "mode_rows", "mode_pct"])
    got, exp = frame_by_attr(df), table_by_attr(nb_stats[17])
    for a, row in exp.items():  # median is a GK-sketch value in Spark: compare by rank distance
        if cell_value(row["median"]) is None:
            assert got[a]["median"] is None or math.isnan(got[a]["median"])
        else:
            _rank_close(O.ColumnProfile(income, a), cell_value(row["median"]), 0.5)
    checked = 0
    for a in exp:
        p = O.ColumnProfile(income, a)
        if p.n == 0:
            continue
        cnt = np.unique(p.nn.astype(str) if p.sdtype == "string" else p.nn, return_counts=True)[1]
        if (cnt == cnt.max()).sum() > 1:
            continue  # tie: the reference's choice is arbitrary (stats_generator.py:358)
        assert str(got[a]["mode"]) == exp[a]["mode"], (a, got[a]["mode"], exp[a]["mode"])
        checked += 1
    assert checked >= 15


def test_nb_cardinality_exact(income, nb_stats):
    _check_table(O.measures_of_cardinality(income, use_approx_unique_count=False), nb_stats[24],
                 ["unique_values", "IDness"])


def test_nb_cardinality_hll_default(income, nb_stats):
    """HLL++ rsd 0.05 (p=9): 20/20 + the all-null column reproduce bit-for-bit."""
    df = O.measures_of_cardinality(income, with_flags=True)
    assert not df["hll_bias_band"].any()
    _check_table(df, nb_stats[23], ["unique_values", "IDness"])


def test_nb_cardinality_hll_rsd002(income, nb_stats):
    """rsd 0.02 (p=12): exact outside the bias-correction band; the band rows are
    flagged 'parity unpinned' (bias tables unavailable offline)."""
    df = O.measures_of_cardinality(income, rsd=0.02, with_flags=True)
    band = set(df.loc[df["hll_bias_band"], "attribute"])
    assert band == {"geohash", "logfnl", "longitude"}
    skip = {(a, c) for a in band for c in ("unique_values", "IDness")}
    _check_table(df, nb_stats[25], ["unique_values", "IDness"], skip=skip)
    got, exp = frame_by_attr(df), table_by_attr(nb_stats[25])
    for a in band:  # exact-distinct fallback is within 3*rsd of Spark's bias-corrected estimate
        assert abs(got[a]["unique_values"] - float(exp[a]["unique_values"])) <= 0.06 * float(exp[a]["unique_values"])


def test_nb_dispersion(income, nb_stats):
    # IQR comes from Spark's GK sketch (rank error <= 1e-4 n): checked by rank below
    _check_table(O.measures_of_dispersion(income), nb_stats[31], ["stddev", "variance", "cov", "range"])
    got, exp = frame_by_attr(O.measures_of_dispersion(income)), table_by_attr(nb_stats[31])
    close = sum(shown_close(got[a]["IQR"], exp[a]["IQR"]) for a in exp)
    assert close >= 6  # IQR = difference of two GK-sketch values; most coincide with the exact ranks


def test_nb_shape(income, nb_stats):
    _check_table(O.measures_of_shape(income), nb_stats[39], ["skewness", "kurtosis"])


def test_nb_percentiles_rank_distance(income, nb_stats):
    """Spark's summary() percentiles are GK-sketch values: each must be a data
    element whose rank is within 1e-4*n of the oracle's exact rank ceil(p*n)."""
    exp = table_by_attr(nb_stats[35])
    df = frame_by_attr(O.measures_of_percentiles(income))
    exact = 0
    for a, row in exp.items():
        p = O.ColumnProfile(income, a)
        assert shown_close(df[a]["min"], row["min"]) and shown_close(df[a]["max"], row["max"])
        for name, q in S.SUMMARY_PCTS.items():
            exact += _rank_close(p, cell_value(row[name]), q) == 0
    assert exact >= 60  # most are the exact element


def test_nb_drift_psi(income, income_source, nb_drift, tmp_path):
    df = O.statistics(income, income_source, source_path=str(tmp_path))
    got, exp = frame_by_attr(df), table_by_attr(nb_drift[6])
    assert set(got) == set(exp)
    for a, row in exp.items():
        g = got[a]["PSI"]
        g = 0.0 if g is None else g
        assert abs(g - float(row["PSI"])) < 5e-7, (a, g, row["PSI"])
        assert got[a]["flagged"] == int(row["flagged"])


def test_nb_drift_other_metrics(income, income_source, nb_drift, tmp_path):
    df = O.statistics(income, income_source, list_of_cols=["age", "education-num", "capital-gain", "hours-per-week"],
                      method_type=["JSD", "HD", "KS"], bin_size=100, source_path=str(tmp_path))
    assert list(df.columns) == ["attribute", "HD", "JSD", "KS", "flagged"]
    assert (df[["HD", "JSD", "KS"]].abs().values < 1e-12).all()


# ---- primitives -------------------------------------------------------------

def test_half_up_and_strings():
    assert S.round_half_up(0.12345) == 0.1235 and S.round_half_up(2.5, 0) == 3.0
    assert S.round_half_up(-0.00005) == -0.0001 and S.round_half_up(None) is None
    assert math.isnan(S.round_half_up(float("nan")))
    assert S.java_double_to_string(5.093362141) == "5.093362141"
    assert S.java_double_to_string(1e7) == "1.0E7" and S.java_double_to_string(1e-4) == "1.0E-4"
    assert S.java_double_to_string(123456.0) == "123456.0" and S.java_double_to_string(0.001) == "0.001"


def test_quantile_rank_float_artefact():
    # 3*(1/10)*10 = 3.0000000000000004 -> rank 4 (pinned by the drift test)
    assert S.quantile_rank(3 * (1 / 10), 10) == 4 and S.quantile_rank(0.5, 4) == 2


def test_xxh64_known_answers():
    # XXH64 reference vectors (seed 0): empty input and "a"
    assert S.xxh64_bytes(b"", 0) == 0xEF46DB3751D8E999
    assert S.xxh64_bytes(b"a", 0) == 0xD24EC4F1A98C6E5B
    assert S.xxh64_bytes(b"abc", 0) == 0x44BC2CF5AD770999
    a = np.array([1, -7, 123456], dtype=np.int32)
    for v, h in zip(a.tolist(), S.xxh64_int_np(a).tolist()):
        assert h == S.xxh64_bytes(int(v).to_bytes(4, "little", signed=True))
    b = np.array([1, -7, 1 << 40], dtype=np.int64)
    for v, h in zip(b.tolist(), S.xxh64_long_np(b).tolist()):
        assert h == S.xxh64_bytes(int(v).to_bytes(8, "little", signed=True))


# ---- N1: drift_stability/test_stability.py:69-92 and notebook cells 16/17 --------------------------------

def _stab_tables():
    l1 = np.array([4.34, 4.76, 4.32, 3.39, 3.67, 4.61, 4.03, 4.93, 3.84, 3.31])
    l2 = np.array([6.34, 4.76, 6.32, 3.39, 5.67, 4.61, 6.03, 4.93, 5.84, 3.31])
    l3 = np.array([8.34, 4.76, 8.32, 3.39, 7.67, 4.61, 8.03, 4.93, 3.84, 3.31])
    return [pa.table({"A": l}) for l in (l1, l2, l3)]


def test_stability_index_known_answers():
    r = O.stability_index_computation(_stab_tables()).iloc[0]
    np.testing.assert_almost_equal([r[c] for c in ("mean_cv", "stddev_cv", "kurtosis_cv", "mean_si", "stddev_si", "kurtosis_si",
                                                   "stability_index", "flagged")], [0.162, 0.62, 0.198, 2.0, 0.0, 2.0, 1.4, 0.0], 3)
    b = [pa.table({"A": np.array([0] * z + [1] * (20 - z))}) for z in (10, 12, 14)]
    r = O.stability_index_computation(b, binary_cols="A").iloc[0]
    np.testing.assert_almost_equal([r["mean_stddev"], r["mean_si"], r["stability_index"], r["flagged"]], [0.1, 0.0, 0.0, 1.0], 3)
    with pytest.raises(ValueError):
        O.stability_index_computation(_stab_tables(), metric_weightages={"mean": 0.5})
    with pytest.raises(ValueError):
        O.stability_index_computation(_stab_tables(), threshold=5)
    with pytest.raises(TypeError):
        O.stability_index_computation(_stab_tables(), binary_cols="Z")


def _stab_datasets():
    import pyarrow.parquet as pq
    import pyarrow.compute as pc
    from conftest import GOLDEN
    t = pq.read_table(GOLDEN + "/stability.parquet")
    return [t.filter(pc.equal(t["_ds"], k)).drop_columns(["_ds"]) for k in range(12)]


def check_stability_notebook(fn, nb_drift, tmp_path):
    """Shared by the oracle and the product test: notebook cells 14-17 of drift_stability.ipynb.
    (The stored stability_index column predates the current compute_si: it shows NaN whenever one
    component score is 0, which the v1.1.0 code no longer does - those cells are skipped.)"""
    ds = _stab_datasets()
    p1, p5, p12 = str(tmp_path / "m1"), str(tmp_path / "m5"), str(tmp_path / "m12")
    r1 = fn([ds[0]], appended_metric_path=p1)
    assert r1["stability_index"].isna().all() and (r1["flagged"] == 1).all()       # one dataset: all null
    r2 = fn(ds[1:5], existing_metric_path=p1, appended_metric_path=p5, threshold=2)
    r3 = fn(ds[5:12], existing_metric_path=p5, appended_metric_path=p12, threshold=2)
    for got, cell in ((r2, 16), (r3, 17)):
        exp = table_by_attr(nb_drift[cell])
        g = frame_by_attr(got)
        assert set(g) == set(exp)
        for a, row in exp.items():
            for c in ("mean_stddev", "mean_cv", "stddev_cv", "kurtosis_cv", "mean_si", "stddev_si", "kurtosis_si"):
                assert shown_close(g[a][c], row[c]), (cell, a, c, g[a][c], row[c])
            if cell_value(row["stability_index"]) is not None:
                assert shown_close(g[a]["stability_index"], row["stability_index"]) and g[a]["flagged"] == int(row["flagged"])


def test_nb_stability_index(nb_drift, tmp_path):
    check_stability_notebook(lambda tables, **kw: O.stability_index_computation(tables, **kw), nb_drift, tmp_path)


# ---- N3: data_analyzer/test_association_evaluator.py:25-240 (IV / IG on the income parquet part-00000) ----------

IV_PINS = {"relationship": 1.6208, "marital-status": 1.3929, "occupation": 0.7686, "education": 0.7525, "education-num": 0.7095,
           "capital-gain": 0.3184, "sex": 0.3111, "workclass": 0.1686}
IG_PINS = {"relationship": 0.1702, "marital-status": 0.1608, "occupation": 0.0916, "education": 0.0938, "education-num": 0.0887,
           "capital-gain": 0.0434, "sex": 0.0379, "workclass": 0.0223}
# these two reference pins come out of approxQuantile(0.01) cutoffs: reproduced by the GK single-batch rank rule
# (spark_semantics.approx_quantile_rank); the textbook rank ceil(p*n) only lands within 10 % of them
GK_PINS = {"age": (1.1891, 0.0944), "hours-per-week": (0.4441, 0.0549)}


def label_table(income_part0):
    inc = np.array(income_part0.column("income").to_pylist(), dtype=object)
    return income_part0.drop_columns(["income"]).append_column("label", pa.array(np.where(inc == "<=50K", 0.0, 1.0)))


def check_iv_ig(iv, ig):
    iv, ig = frame_by_attr(iv), frame_by_attr(ig)
    assert len(iv) == 15 and len(ig) == 15
    for a, v in IV_PINS.items():
        assert round(iv[a]["iv"], 4) == v, (a, iv[a]["iv"], v)
    for a, v in IG_PINS.items():
        assert round(ig[a]["ig"], 4) == v, (a, ig[a]["ig"], v)
    for a, (v1, v2) in GK_PINS.items():
        assert round(iv[a]["iv"], 4) == v1 and round(ig[a]["ig"], 4) == v2, (a, iv[a]["iv"], ig[a]["ig"])


def test_iv_ig_known_answers(income_part0):
    t = label_table(income_part0)
    check_iv_ig(O.IV_calculation(t, drop_cols=["ifa"]), O.IG_calculation(t, drop_cols=["ifa"]))
    with pytest.raises(TypeError):
        O.IV_calculation(t, label_col="nope")
    with pytest.raises(TypeError):
        O.IV_calculation(t, event_label=7)


# ---- N2: outlier_detection (data_analyzer/test_quality_checker.py:526-668) and the GK rank rule -------------------

OUTLIER_PINS_UPPER = {"age": [0, 87, 0], "fnlwgt": [0, 518, 0], "logfnl": [0, 15, 0], "education-num": [0, 0, 0],
                      "capital-gain": [0, 955, 0], "hours-per-week": [0, 515, 0], "capital-loss": [0, 0, 1]}   # :548-554


def check_outlier_reference_tests(run, table, tmp_path):
    """The reference's four outlier tests; `run(table, **kw)` -> (odf, odf_print pandas) hides oracle vs product."""
    n_rows, n_cols = table.num_rows, len(table.column_names)
    odf, pr = run(table, drop_cols=["ifa", "label"], treatment=True, treatment_method="row_removal", print_impact=True)   # :526-555
    assert n_rows > odf["rows"] and odf["columns"] == table.column_names and pr.shape == (7, 4)
    got = {r["attribute"]: [r["lower_outliers"], r["upper_outliers"], r["excluded_due_to_skewness"]] for r in pr.to_dict("records")}
    assert got == OUTLIER_PINS_UPPER
    odf, pr = run(table, list_of_cols=["age", "education-num"], detection_side="both",                                  # :558-592
                  detection_configs={"pctile_lower": 0.02, "pctile_upper": 0.98}, treatment=True, output_mode="append",
                  print_impact=True)
    assert odf["rows"] == n_rows and len(odf["columns"]) == n_cols + 2
    assert odf["minmax"]("age") == (17, 85) and odf["minmax"]("age_outliered") == (18, 66)
    assert odf["minmax"]("education-num") == (1, 16) and odf["minmax"]("education-num_outliered") == (4, 15)
    got = {r["attribute"]: [r["lower_outliers"], r["upper_outliers"], r["excluded_due_to_skewness"]] for r in pr.to_dict("records")}
    assert got == {"age": [202, 482, 0], "education-num": [267, 205, 0]}
    model = str(tmp_path / "outlier_model")                                                                              # :595-637
    same = run(table, list_of_cols=["logfnl", "hours-per-week", "capital-loss"], detection_side="both", treatment=False,
               model_path=model, print_impact=False)
    assert same["rows"] == n_rows and same["columns"] == table.column_names
    odf, pr = run(table, list_of_cols=["logfnl", "hours-per-week", "capital-gain", "capital-loss"], detection_side="lower",
                  treatment=True, treatment_method="null_replacement", pre_existing_model=True, model_path=model,
                  print_impact=True)
    assert odf["rows"] == n_rows and odf["columns"] == table.column_names and pr.shape == (3, 4)
    assert odf["nulls"]("hours-per-week") == table.column("hours-per-week").null_count + 825
    assert odf["nulls"]("logfnl") == table.column("logfnl").null_count + 314
    got = {r["attribute"]: [r["lower_outliers"], r["upper_outliers"], r["excluded_due_to_skewness"]] for r in pr.to_dict("records")}
    assert got == {"hours-per-week": [825, 0, 0], "logfnl": [314, 0, 0], "capital-loss": [0, 0, 1]}
    with pytest.raises(TypeError):                                                                                       # :640-656
        run(table, list_of_cols=["capital-gain", "capital-loss"], detection_side="both",
            detection_configs={"pctile_lower": 0.05, "stdev_lower": 3.0, "stdev_upper": 3.0}, treatment=True)
    with pytest.raises(TypeError):                                                                                       # :658-668
        run(table, list_of_cols=["capital-gain", "capital-loss"], detection_side="both",
            detection_configs={"stdev_lower": 3.0, "stdev_upper": 3.0, "min_validation": 3}, treatment=True)


def _oracle_outlier_run(tmp_models={}):
    import pyarrow.compute as pc

    def run(table, print_impact=False, model_path="NA", pre_existing_model=False, **kw):
        params = None
        cfg = kw.get("detection_configs")
        if pre_existing_model:
            kept, bounds, skewed = tmp_models[model_path]
            cols = kw["list_of_cols"]
            params = ([c for c in kept if c in cols], [b for c, b in zip(kept, bounds) if c in cols], [c for c in skewed if c in cols])
        elif model_path != "NA":
            cols = kw["list_of_cols"]
            tmp_models[model_path] = O.outlier_bounds(table, cols, kw.get("detection_side", "upper"),
                                                      O._DEFAULT_OUTLIER_CFG if cfg is None else cfg)
        O.outlier_methodologies(kw.get("detection_side", "upper"), O._DEFAULT_OUTLIER_CFG if cfg is None else cfg)
        odf, pr = O.outlier_detection(table, params=params, **kw)
        view = {"rows": odf.num_rows, "columns": odf.column_names,
                "minmax": lambda c: (pc.min(odf[c]).as_py(), pc.max(odf[c]).as_py()),
                "nulls": lambda c: odf.column(c).null_count}
        return (view, pr) if print_impact else view
    return run


def test_outlier_detection_known_answers(income_part0, tmp_path):
    check_outlier_reference_tests(_oracle_outlier_run(), income_part0.append_column("label", pa.array([0] * income_part0.num_rows)),
                                  tmp_path)


def test_gk_rank_rule():
    """Spark's sketch for one partition of < 50 000 values: data-independent positions.  Small n: no compression ->
    the exact rank (median([23,42,51,55]) = 42, test_stats_generator.py:292-339); the product's closed form equals
    the restated loop; beyond the head buffer the exact rule applies."""
    from anovos_b200.shared import gk
    assert S.approx_quantile_rank(0.5, 4, 1e-4) == 2 and S.approx_quantile_rank(0.5, 4, 0.01) == 2
    for n in (1, 2, 3, 7, 49, 50, 51, 100, 101, 999, 5000, 5001, 6102, 16250, 32561, 49999):
        for eps in (1e-4, 0.01, 0.05):
            sm = S.gk_single_batch_summary(n, eps)
            assert sm[0][0] == 0 and sm[-1][0] == n - 1 and sum(g for _, g, _ in sm) == n
            for p in (0.0, 0.01, 0.02, 0.05, 0.1, 0.25, 3 * 0.1, 0.5, 7 * (1 / 10), 0.75, 0.95, 0.98, 0.99, 1.0):
                r = S.gk_query_position(sm, n, eps, p) + 1
                assert r == gk.spark_rank(n, p, eps), (n, eps, p)
                assert abs(r - S.quantile_rank(p, n)) <= max(1, math.ceil(eps * n) + 1), (n, eps, p, r)   # inside Spark's band
                if 2 * eps * n <= 1:
                    assert r == S.quantile_rank(p, n)
    assert gk.spark_rank(50000, 0.5, 0.01) == 25000 and gk.spark_rank(0, 0.5, 0.01) == 0
    assert S.approx_quantile_rank(0.5, 60000, 0.01) == 30000


def test_nb_percentiles_exact_with_spark_partitions(income, nb_stats):
    """With the partitioning Spark used for the notebook run (two Hadoop splits of the CSV at 4 MiB,
    tests/golden/income_partitions.json) the restated sketch - one sorted batch per partition, compress, merge in
    partition order, query - reproduces ALL 81 stored summary() percentiles, every median and every IQR."""
    import json
    import os
    from conftest import GOLDEN
    parts = json.load(open(os.path.join(GOLDEN, "income_partitions.json")))["rows_per_partition"]
    assert sum(parts) == income.num_rows and len(parts) == 2
    t = O.with_spark_partitions(income, parts)
    _check_table(O.measures_of_percentiles(t), nb_stats[35], ["min"] + list(S.SUMMARY_PCTS) + ["max"])
    _check_table(O.measures_of_centralTendency(t), nb_stats[17], ["mean", "median", "mode_rows", "mode_pct"])
    _check_table(O.measures_of_dispersion(t), nb_stats[31], ["stddev", "variance", "cov", "IQR", "range"])


# ---- the association notebook: IV / IG on the full income CSV (two Spark partitions) --------------------------------------

NB_IV_IG_CALLS = {  # code cell -> kwargs of the call stored in the notebook (label_col="income", event_label=">50K")
    18: {}, 19: {"drop_cols": ["ifa"]}, 20: {"list_of_cols": ["age", "sex", "race", "workclass", "fnlwgt"]},
    21: {"list_of_cols": ["age", "sex", "race", "workclass", "fnlwgt"],
         "encoding_configs": {"bin_method": "equal_range", "bin_size": 10, "monotonicity_check": 0}},
    22: {"list_of_cols": ["age", "sex", "race", "workclass", "fnlwgt"],
         "encoding_configs": {"bin_method": "equal_frequency", "bin_size": 20, "monotonicity_check": 0}},
}


def check_nb_iv_ig(iv_fn, ig_fn, nb_assoc):
    """All stored IV (cells 18-22) and IG (cells 25-29) tables: 54 + 54 values, to the 6 decimals shown.  The numeric
    attributes hang on approxQuantile(0.01) cutoffs over two merged partition sketches; ifa (an id) has a NULL ig."""
    n = 0
    for cell, kw in NB_IV_IG_CALLS.items():
        for fn, c, col in ((iv_fn, cell, "iv"), (ig_fn, cell + 7, "ig")):
            got, exp = frame_by_attr(fn(label_col="income", event_label=">50K", **kw)), table_by_attr(nb_assoc[c])
            assert set(got) == set(exp), (c, sorted(set(got) ^ set(exp)))
            for a, row in exp.items():
                assert shown_close(None if pd.isna(got[a][col]) else got[a][col], row[col]), (c, a, got[a][col], row[col])
                n += 1
    assert n == 108


def test_nb_iv_ig_with_spark_partitions(income_spark, nb_assoc):
    import functools
    check_nb_iv_ig(functools.partial(O.IV_calculation, income_spark), functools.partial(O.IG_calculation, income_spark), nb_assoc)


# ---- the transformers notebook: attribute_binning on the income CSV, first five rows as displayed (cells 6 and 8) -------------

NB_BINNING = {   # (education-num, hours-per-week) bin ids of rows 1a..5a
    "equal_range": [[4, 3], [4, 1], [3, 3], [2, 3], [4, 3]],          # cell 6: bin_size=5, output_mode="append"
    "equal_frequency": [[4, 2], [4, 1], [1, 2], [1, 2], [4, 2]],      # cell 8: bin_size=5 (approxQuantile over two partitions)
}


def test_nb_attribute_binning_head(income_spark):
    for method, exp in NB_BINNING.items():
        out = O.attribute_binning(income_spark, list_of_cols=["education-num", "hours-per-week"], method_type=method, bin_size=5)
        assert out.column("ifa").slice(0, 5).to_pylist() == ["1a", "2a", "3a", "4a", "5a"]
        got = [[out.column(c)[i].as_py() for c in ("education-num", "hours-per-week")] for i in range(5)]
        assert got == exp, (method, got)
