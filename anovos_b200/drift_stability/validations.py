"""Argument normalisation decorators of `drift_detector.statistics` (reference
/root/reference/src/main/anovos/drift_stability/validations.py:8-94).  Like the reference
they read KEYWORD arguments only: a positional list_of_cols is ignored (SURVEY C#6)."""
from __future__ import annotations

from functools import partial, wraps

from ..frame import as_frame
from ..shared.utils import attributeType_segregation


def check_list_of_columns(func=None, columns="list_of_cols", target_idx: int = 1, target: str = "idf_target",
                          drop="drop_cols"):
    if func is None:
        return partial(check_list_of_columns, columns=columns, target=target, drop=drop)

    @wraps(func)
    def validate(*args, **kwargs):
        idf_target = kwargs.get(target, None)
        if idf_target is None:
            idf_target = args[target_idx]
        idf_target = as_frame(idf_target)

        cols_raw = kwargs.get(columns, "all")
        if isinstance(cols_raw, str):
            if cols_raw == "all":
                num_cols, cat_cols, _ = attributeType_segregation(idf_target)
                cols = num_cols + cat_cols
            else:
                cols = [x.strip() for x in cols_raw.split("|")]
        elif isinstance(cols_raw, list):
            cols = cols_raw
        else:
            raise TypeError(f"'{columns}' must be either a string or a list of strings. Received {type(cols_raw)}.")

        drops_raw = kwargs.get(drop, [])
        if drops_raw is None:
            drops_raw = []
        if isinstance(drops_raw, str):
            drops = [x.strip() for x in drops_raw.split("|")]
        elif isinstance(drops_raw, list):
            drops = drops_raw
        else:
            raise TypeError(f"'{drop}' must be either a string or a list of strings. Received {type(drops_raw)}.")

        final_cols = []
        for e in cols:
            if e not in drops and e not in final_cols:
                final_cols.append(e)
        if not final_cols:
            raise ValueError(f"Empty set of columns is given. Columns to select: {cols}, columns to drop: {drops}.")
        if any(x not in idf_target.columns for x in final_cols):
            raise ValueError("Not all columns are in the input dataframe. "
                             f"Missing columns: {set(final_cols) - set(idf_target.columns)}")
        kwargs[columns] = final_cols
        kwargs[drop] = []
        return func(*args, **kwargs)

    return validate


def check_distance_method(func=None, param="method_type"):
    if func is None:
        return partial(check_distance_method, param=param)

    @wraps(func)
    def validate(*args, **kwargs):
        methods = kwargs.get(param, "PSI")
        if isinstance(methods, str):
            methods = ["PSI", "JSD", "HD", "KS"] if methods == "all" else [x.strip() for x in methods.split("|")]
        if any(x not in ("PSI", "JSD", "HD", "KS") for x in methods):
            raise TypeError(f"Invalid input for {param}")
        kwargs[param] = methods
        return func(*args, **kwargs)

    return validate
