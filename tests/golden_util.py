"""Helpers to compare result frames with the notebook tables in tests/golden/."""
import math


def cell_value(s):
    if s in ("NaN", "None", ""):
        return None
    try:
        return float(s)
    except ValueError:
        return s


def shown_close(got, shown: str, decimals_hint=4):
    """`shown` is a pandas-rendered number (6 significant digits or fixed 4/6 decimals)."""
    exp = cell_value(shown)
    if exp is None:
        return got is None or (isinstance(got, float) and math.isnan(got))
    if isinstance(exp, str):
        return str(got) == exp
    if got is None:
        return False
    got = float(got)
    if "e" in shown.lower():
        return abs(got - exp) <= 1e-6 * abs(exp) + 1e-12
    nd = len(shown.split(".")[1]) if "." in shown else 0
    return abs(got - exp) <= 0.5000001 * 10 ** (-nd) + 1e-12


def frame_by_attr(df):
    return {r["attribute"]: r for r in df.to_dict("records")}


def table_by_attr(t):
    cols = t["columns"]
    return {r[0]: dict(zip(cols, r)) for r in t["rows"]}
