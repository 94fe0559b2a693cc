"""Result object of the B200 path: the small per-attribute frames the reference returns as
Spark DataFrames.  Callers of the reference immediately do `.toPandas().to_csv(...)`
(data_report/report_preprocessing.py:92), `.show(n)` (workflow.py:509), `.count()` or
`.where(F.col("attribute") == x)` (reference tests); this class offers that surface on top
of a pandas DataFrame."""
from __future__ import annotations

import pandas as pd


class ResultFrame:
    def __init__(self, df: pd.DataFrame):
        idx = df.index
        plain = isinstance(idx, pd.RangeIndex) and idx.start == 0 and idx.step == 1
        self._df = df if plain else df.reset_index(drop=True)

    def toPandas(self) -> pd.DataFrame:
        return self._df.copy()

    to_pandas = toPandas

    @property
    def columns(self):
        return list(self._df.columns)

    def count(self) -> int:
        return len(self._df)

    def show(self, n: int = 20, truncate=True):
        with pd.option_context("display.max_columns", None, "display.width", 200):
            print(self._df.head(n).to_string(index=False))

    def where(self, cond) -> "ResultFrame":
        """cond: pandas query string ("attribute == 'age'"), a callable df -> mask, or a
        dict {column: value}."""
        if isinstance(cond, str):
            return ResultFrame(self._df.query(cond))
        if isinstance(cond, dict):
            m = pd.Series(True, index=self._df.index)
            for k, v in cond.items():
                m &= self._df[k] == v
            return ResultFrame(self._df[m])
        return ResultFrame(self._df[cond(self._df)])

    filter = where

    def select(self, *cols) -> "ResultFrame":
        cols = list(cols[0]) if len(cols) == 1 and isinstance(cols[0], (list, tuple)) else list(cols)
        return ResultFrame(self._df[cols])

    def collect(self):
        return [tuple(r) for r in self._df.itertuples(index=False)]

    def to_csv(self, path, **kw):
        kw.setdefault("index", False)
        return self._df.to_csv(path, **kw)

    def __len__(self):
        return len(self._df)

    def __repr__(self):
        return "ResultFrame(%d rows x %d cols: %s)" % (len(self._df), self._df.shape[1], ", ".join(self.columns))
