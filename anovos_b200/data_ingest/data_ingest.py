"""`read_dataset` / `write_dataset` of `anovos.data_ingest.data_ingest` (reference
/root/reference/src/main/anovos/data_ingest/data_ingest.py:23-110) for the formats either side of the
hot path: csv and parquet files or Spark-style part-file directories go straight into a ColumnFrame
(pyarrow parses on the host; the Arrow buffers are uploaded as they are, no row-wise conversion)
and result / frame objects are written back in the layouts the reference produces."""
from __future__ import annotations

import os

from ..frame import ColumnFrame, as_frame
from ..result import ResultFrame


def _truthy(v):
    return str(v).lower() == "true"


def _part_files(path, ext):
    if os.path.isdir(path):
        files = sorted(os.path.join(path, f) for f in os.listdir(path)
                       if f.endswith("." + ext) or (ext == "parquet" and f.endswith(".snappy.parquet")))
        if not files:
            raise FileNotFoundError("no .%s part files under %s" % (ext, path))
        return files
    return [path]


def read_dataset(spark, file_path, file_type, file_configs={}):
    """-> ColumnFrame.  csv options honoured: header, delimiter / sep, inferSchema (without it every column is
    a string, like Spark), nullValue; empty fields are nulls (Spark's default)."""
    import pyarrow as pa
    import pyarrow.csv as pacsv
    import pyarrow.parquet as pq
    if file_type == "parquet":
        t = pa.concat_tables([pq.read_table(f) for f in _part_files(file_path, "parquet")], promote_options="default")
    elif file_type == "csv":
        cfg = {k.lower(): v for k, v in file_configs.items()}
        header = _truthy(cfg.get("header", "false"))
        delim = cfg.get("delimiter", cfg.get("sep", ","))
        infer = _truthy(cfg.get("inferschema", "false"))
        null_values = [cfg.get("nullvalue", "")]
        tables = []
        for f in _part_files(file_path, "csv"):
            ropt = pacsv.ReadOptions(autogenerate_column_names=not header)
            popt = pacsv.ParseOptions(delimiter=delim)
            if infer:
                copt = pacsv.ConvertOptions(strings_can_be_null=True, null_values=null_values)
                tb = pacsv.read_csv(f, read_options=ropt, parse_options=popt, convert_options=copt)
            else:
                names = pacsv.open_csv(f, read_options=ropt, parse_options=popt).schema.names
                copt = pacsv.ConvertOptions(column_types={n: pa.string() for n in names}, strings_can_be_null=True,
                                            null_values=null_values)
                tb = pacsv.read_csv(f, read_options=ropt, parse_options=popt, convert_options=copt)
            if not header:
                tb = tb.rename_columns(["_c%d" % i for i in range(tb.num_columns)])   # Spark's default names
            tables.append(tb)
        t = pa.concat_tables(tables, promote_options="default")
        if infer:  # Spark infers IntegerType for integers that fit 32 bits and StringType for all-null columns
            import pyarrow.compute as pc
            for i, fld in enumerate(t.schema):
                if pa.types.is_null(fld.type):
                    t = t.set_column(i, fld.name, t.column(i).cast(pa.string()))
                    continue
                if pa.types.is_int64(fld.type) and t.num_rows:
                    mm = pc.min_max(t.column(i))
                    lo, hi = mm["min"].as_py(), mm["max"].as_py()
                    if lo is None or (-(1 << 31) <= lo and hi < (1 << 31)):
                        t = t.set_column(i, fld.name, t.column(i).cast(pa.int32()))
    else:
        raise NotImplementedError("file_type %r: only csv and parquet are part of the B200 hot-path build" % file_type)
    return ColumnFrame.from_arrow(t)


def write_dataset(idf, file_path, file_type, file_configs={}, column_order=[]):
    """Writes a ResultFrame / pandas / pyarrow object as `<file_path>/part-00000.<ext>` (Spark-style directory).
    mode: error (default) | overwrite."""
    import pandas as pd
    import pyarrow as pa
    import pyarrow.parquet as pq
    if isinstance(idf, ResultFrame):
        df = idf.toPandas()
    elif isinstance(idf, pd.DataFrame):
        df = idf
    elif isinstance(idf, pa.Table):
        df = idf.to_pandas()
    else:
        raise TypeError("write_dataset: pass a ResultFrame, pandas DataFrame or pyarrow Table")
    if column_order:
        if len(column_order) != len(df.columns):
            raise ValueError("Count of column(s) specified in column_order argument do not match Dataframe")
        diff = [c for c in column_order if c not in set(df.columns)]
        if diff:
            raise ValueError("Column(s) specified in column_order argument not found in Dataframe: " + str(diff))
        df = df[list(column_order)]
    mode = file_configs.get("mode", "error")
    if os.path.exists(file_path):
        if mode == "error":
            raise FileExistsError(file_path)
        for f in os.listdir(file_path):
            os.remove(os.path.join(file_path, f))
    os.makedirs(file_path, exist_ok=True)
    if file_type == "csv":
        df.to_csv(os.path.join(file_path, "part-00000.csv"), index=False, header=_truthy(file_configs.get("header", "false")),
                  sep=file_configs.get("delimiter", ","))
    elif file_type == "parquet":
        pq.write_table(pa.Table.from_pandas(df, preserve_index=False), os.path.join(file_path, "part-00000.parquet"))
    else:
        raise NotImplementedError("file_type %r" % file_type)
    open(os.path.join(file_path, "_SUCCESS"), "w").close()
