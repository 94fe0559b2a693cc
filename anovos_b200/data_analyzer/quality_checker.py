"""Consumers of the stats hot path from `anovos.data_analyzer.quality_checker` (SURVEY.md 8f, row
N2), same signatures and outputs as the reference
(/root/reference/src/main/anovos/data_analyzer/quality_checker.py):
  nullColumns_detection :286-547  (treatment: none / row_removal / column_removal)
  IDness_detection      :1048-1182
  biasedness_detection  :1185-1339
Each returns (odf, odf_print).  The per-row work (null counts, distinct counts, modes) runs in
the CUDA kernels through stats_generator.  outlier_detection, duplicate / invalid-entry detection
and the imputation treatments (MMM / KNN / regression / MF / auto) are not part of this build."""
from __future__ import annotations

import warnings

import pandas as pd

from .. import profile
from ..frame import as_frame
from ..result import ResultFrame
from ..shared.utils import attributeType_segregation, spark_round
from .stats_generator import _mode_str, measures_of_cardinality, missingCount_computation


def _names(x):
    if isinstance(x, str):
        return [s.strip() for s in x.split("|")]
    return list(x)


def _as_bool(v, what):
    if str(v).lower() == "true":
        return True
    if str(v).lower() == "false":
        return False
    raise TypeError("Non-Boolean input for " + what)


def _unique(cols, drop):
    out = []
    for c in cols:
        if c not in drop and c not in out:
            out.append(c)
    return out


def _read_stats(spec, columns):
    """read_dataset(spark, **stats_x): file_path + file_type (csv / parquet) of a saved stats frame."""
    path, ftype = spec["file_path"], spec.get("file_type", "csv")
    import os
    if os.path.isdir(path):
        files = sorted(os.path.join(path, f) for f in os.listdir(path) if f.endswith("." + ftype))
    else:
        files = [path]
    df = pd.concat([pd.read_csv(f) if ftype == "csv" else pd.read_parquet(f) for f in files], ignore_index=True)
    return df[columns]


def nullColumns_detection(spark, idf, list_of_cols="missing", drop_cols=[], treatment=False, treatment_method="row_removal",
                          treatment_configs={}, stats_missing={}, stats_unique={}, stats_mode={}, print_impact=False):
    fr = as_frame(idf)
    if stats_missing == {}:
        stats = missingCount_computation(spark, fr).toPandas()
    else:
        stats = _read_stats(stats_missing, ["attribute", "missing_count", "missing_pct"])
    missing_cols = stats.loc[stats["missing_count"] > 0, "attribute"].tolist()
    if isinstance(list_of_cols, str) and list_of_cols == "all":
        num, cat, _ = attributeType_segregation(fr)
        list_of_cols = num + cat
    if isinstance(list_of_cols, str) and list_of_cols == "missing":
        list_of_cols = missing_cols
    cols = _unique(_names(list_of_cols), _names(drop_cols))
    if not cols:
        warnings.warn("No Null Detection - No column(s) to analyze")
        return fr, ResultFrame(pd.DataFrame(columns=["attribute", "missing_count", "missing_pct"]))
    if any(c not in fr.columns for c in cols):
        raise TypeError("Invalid input for Column(s)")
    treatment = _as_bool(treatment, "treatment")
    if treatment_method not in ("MMM", "row_removal", "column_removal", "KNN", "regression", "MF", "auto"):
        raise TypeError("Invalid input for method_type")
    treatment_configs = dict(treatment_configs)
    threshold = treatment_configs.pop("treatment_threshold", None)
    if threshold:
        threshold = float(threshold)
    elif treatment_method == "column_removal":
        raise TypeError("Invalid input for column removal threshold")
    stats = stats[stats["attribute"].isin(cols)].reset_index(drop=True)
    odf = fr
    if treatment:
        threshold_cols = stats.loc[stats["missing_pct"] > threshold, "attribute"].tolist() if threshold else []
        if treatment_method == "column_removal":
            odf = fr.drop(threshold_cols)
            if print_impact:
                print("Removed Columns: ", threshold_cols)
        elif treatment_method == "row_removal":
            remove = stats.loc[stats["missing_pct"] == 1.0, "attribute"].tolist()
            sub = [c for c in cols if c not in remove]
            if threshold:
                sub = [c for c in threshold_cols if c not in remove]
            odf = fr.dropna(subset=sub)
            if print_impact:
                print("Before Count: " + str(fr.count()))
                print("After Count: " + str(odf.count()))
        else:
            raise NotImplementedError("treatment_method=%r (imputation) is outside the B200 hot-path build" % treatment_method)
    out = ResultFrame(stats)
    if print_impact:
        out.show(len(cols))
    return odf, out


def _discrete(fr, list_of_cols, drop_cols):
    if isinstance(list_of_cols, str) and list_of_cols == "all":
        num, cat, _ = attributeType_segregation(fr)
        list_of_cols = num + cat
    cols = _unique(_names(list_of_cols), _names(drop_cols))
    if any(c not in fr.columns for c in cols):
        raise TypeError("Invalid input for Column(s)")
    types = dict(fr.dtypes)
    return [c for c in cols if types[c] in ("string", "int", "bigint", "long")]   # reference :1122-1124


def IDness_detection(spark, idf, list_of_cols="all", drop_cols=[], treatment=False, treatment_threshold=0.8, stats_unique={},
                     print_impact=False):
    fr = as_frame(idf)
    cols = _discrete(fr, list_of_cols, drop_cols)
    if not cols:
        warnings.warn("No IDness Check - No discrete column(s) to analyze")
        return fr, ResultFrame(pd.DataFrame(columns=["attribute", "unique_values", "IDness", "flagged"]))
    treatment_threshold = float(treatment_threshold)
    if treatment_threshold < 0 or treatment_threshold > 1:
        raise TypeError("Invalid input for Treatment Threshold Value")
    treatment = _as_bool(treatment, "treatment")
    if stats_unique == {}:
        stats = measures_of_cardinality(spark, fr, cols).toPandas()
    else:
        stats = _read_stats(stats_unique, ["attribute", "unique_values", "IDness"])
        stats = stats[stats["attribute"].isin(cols)].reset_index(drop=True)
    stats["flagged"] = (stats["IDness"] >= treatment_threshold).astype(int)
    odf = fr
    if treatment:
        remove = stats.loc[stats["flagged"] == 1, "attribute"].tolist()
        odf = fr.drop(remove)
        stats = stats.rename(columns={"flagged": "treated"})
        if print_impact:
            print("Removed Columns: ", remove)
    out = ResultFrame(stats)
    if print_impact:
        out.show(len(cols))
    return odf, out


def biasedness_detection(spark, idf, list_of_cols="all", drop_cols=[], treatment=False, treatment_threshold=0.8, stats_mode={},
                         print_impact=False):
    fr = as_frame(idf)
    cols = _discrete(fr, list_of_cols, drop_cols)
    if not cols:
        warnings.warn("No biasedness Check - No discrete column(s) to analyze")
        return fr, ResultFrame(pd.DataFrame(columns=["attribute", "mode", "mode_rows", "mode_pct", "flagged"]))
    if treatment_threshold < 0 or treatment_threshold > 1:
        raise TypeError("Invalid input for Treatment Threshold Value")
    treatment = _as_bool(treatment, "treatment")
    if stats_mode == {}:
        nv = profile.n_valid(fr, cols)
        md = profile.mode_distinct(fr, cols)
        rows = []
        for c in cols:
            mode, mrows = md[c][0], md[c][1]
            rows.append([c, _mode_str(fr.column(c), mode), mrows, None if mrows is None else spark_round(mrows / nv[c])])
        stats = pd.DataFrame(rows, columns=["attribute", "mode", "mode_rows", "mode_pct"])
    else:
        stats = _read_stats(stats_mode, ["attribute", "mode", "mode_rows", "mode_pct"])
        stats = stats[stats["attribute"].isin(cols)].reset_index(drop=True)
    flag = [(1 if (p is None or pd.isna(p) or p >= treatment_threshold) else 0) for p in stats["mode_pct"].tolist()]
    stats["flagged"] = flag
    odf = fr
    if treatment:
        remove = stats.loc[stats["flagged"] == 1, "attribute"].tolist()
        odf = fr.drop(remove)
        stats = stats.rename(columns={"flagged": "treated"})
        if print_impact:
            print("Removed Columns: ", remove)
    out = ResultFrame(stats)
    if print_impact:
        out.show(len(cols))
    return odf, out
