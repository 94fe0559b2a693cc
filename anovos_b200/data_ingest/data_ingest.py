"""`read_dataset` / `write_dataset` of `anovos.data_ingest.data_ingest` (reference
/root/reference/src/main/anovos/data_ingest/data_ingest.py:23-110) for the formats either side of the
hot path: csv and parquet files or Spark-style part-file directories go straight into a ColumnFrame
(pyarrow parses on the host; the Arrow buffers are uploaded as they are, no row-wise conversion)
and result / frame objects are written back in the layouts the reference produces."""
from __future__ import annotations

import os

import numpy as np

from ..frame import ColumnFrame
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


OPEN_COST = 4 * 1024 * 1024          # spark.sql.files.openCostInBytes
MAX_PARTITION = 128 * 1024 * 1024    # spark.sql.files.maxPartitionBytes


def spark_file_partitions(files, cores):
    """How Spark (FilePartition.getFilePartitions, local[cores]) cuts `files` into scan partitions:
    maxSplitBytes = min(maxPartitionBytes, max(openCostInBytes, sum(len + openCost) / cores)); files by size
    descending, each cut every maxSplitBytes, the pieces packed next-fit.  -> [[(path, start, length), ...], ...]"""
    sized = sorted(((os.path.getsize(f), f) for f in files), key=lambda t: -t[0])
    total = sum(n + OPEN_COST for n, _ in sized)
    max_split = min(MAX_PARTITION, max(OPEN_COST, total // max(int(cores), 1)))
    pieces = [(f, off, min(max_split, n - off)) for n, f in sized for off in range(0, max(n, 1), max_split)]
    parts, cur, size = [], [], 0
    for p in pieces:
        if cur and size + p[2] > max_split:
            parts.append(cur)
            cur, size = [], 0
        cur.append(p)
        size += p[2] + OPEN_COST
    if cur:
        parts.append(cur)
    return parts


def csv_rows_in_range(path, start, length, header):
    """Rows Hadoop's LineRecordReader hands to the split [start, start+length): every line that STARTS at an offset
    <= the split's end, minus the lines of earlier splits (a split with start > 0 skips its first, partial line)."""
    raw = open(path, "rb").read()
    starts = np.flatnonzero(np.frombuffer(raw, np.uint8) == 10) + 1
    starts = np.concatenate([[0], starts[starts < len(raw)]])
    if header:
        starts = starts[1:]

    def upto(end):      # lines starting at an offset <= end
        return int(np.searchsorted(starts, end, "right"))
    lo = 0 if start == 0 else upto(start)
    return upto(start + length) - lo if start + length < len(raw) else len(starts) - lo


def read_dataset(spark, file_path, file_type, file_configs={}):
    """-> ColumnFrame.  csv options honoured: header, delimiter / sep, inferSchema (without it every column is
    a string, like Spark), nullValue; empty fields are nulls (Spark's default).
    Extra option `spark_cores` (N of the `local[N]` being replaced): the frame is returned as the scan partitions
    Spark would create for these files (PartitionedFrame, rows in Spark's partition order), so that summary() /
    approxQuantile percentiles come out exactly as Spark's per-partition sketches would give them."""
    import pyarrow as pa
    import pyarrow.csv as pacsv
    import pyarrow.parquet as pq
    cores = {k.lower(): v for k, v in file_configs.items()}.get("spark_cores")
    per_file = {}
    if file_type == "parquet":
        for f in _part_files(file_path, "parquet"):
            per_file[f] = pq.read_table(f)
        t = pa.concat_tables(list(per_file.values()), promote_options="default")
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
            per_file[f] = tb
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
    elif file_type == "json":
        # Spark's json source reads JSON Lines; inferred struct fields come out sorted by name (JsonInferSchema),
        # integers as bigint, fractions as double
        import pyarrow.json as pajson
        for f in _part_files(file_path, "json"):
            per_file[f] = pajson.read_json(f)
        t = pa.concat_tables(list(per_file.values()), promote_options="default")
        t = t.select(sorted(t.column_names))
        per_file = {f: tb.select([c for c in sorted(t.column_names) if c in tb.column_names]) for f, tb in per_file.items()}
    elif file_type == "avro":
        from .avro_io import read_avro
        for f in _part_files(file_path, "avro"):
            per_file[f] = read_avro(f)
        t = pa.concat_tables(list(per_file.values()), promote_options="default")
    else:
        raise NotImplementedError("file_type %r: csv, parquet, json and avro are supported" % file_type)
    if cores:
        return _as_spark_partitions(t, per_file, int(cores), file_type, file_type == "csv" and header)
    return ColumnFrame.from_arrow(t)


def _as_spark_partitions(t, per_file, cores, file_type, header):
    """Re-order the rows into Spark's scan-partition order and tag the partition sizes."""
    import pyarrow as pa
    from ..partitioned import PartitionedFrame
    files = list(per_file)
    first_row, r = {}, 0
    for f in files:                                   # rows of file f inside t (typed like t)
        first_row[f] = r
        r += per_file[f].num_rows
    consumed = {f: 0 for f in files}
    slices, rows = [], []
    for part in spark_file_partitions(files, cores):
        k_part = 0
        for f, start, length in part:
            if file_type == "csv":
                k = csv_rows_in_range(f, start, length, header)
            elif start == 0 and length >= os.path.getsize(f):
                k = per_file[f].num_rows
            else:
                raise NotImplementedError("spark_cores: parquet files larger than one split are not supported")
            slices.append(t.slice(first_row[f] + consumed[f], k))
            consumed[f] += k
            k_part += k
        rows.append(k_part)
    if any(consumed[f] != per_file[f].num_rows for f in files):
        raise ValueError("spark_cores: split arithmetic does not cover every row (quoted newlines are not supported)")
    keep = [k for k in rows if k]
    return PartitionedFrame.from_arrow_partitions(pa.concat_tables(slices), keep if keep else [0])


def write_dataset(idf, file_path, file_type, file_configs={}, column_order=[]):
    """Writes a ColumnFrame / PartitionedFrame (D2H of values + validity, dictionary decode), ResultFrame, pandas or
    pyarrow object as `<file_path>/part-00000.<ext>` (Spark-style directory + _SUCCESS).  file_type: csv | parquet |
    json | avro; file_configs: header, delimiter, compression (avro: uncompressed | deflate | snappy),
    mode: error (default) | overwrite | append.  Reference data_ingest.py:54-110."""
    import shutil
    import pandas as pd
    import pyarrow as pa
    import pyarrow.parquet as pq
    if isinstance(idf, ResultFrame):
        tb = pa.Table.from_pandas(idf.toPandas(), preserve_index=False)
    elif isinstance(idf, pd.DataFrame):
        tb = pa.Table.from_pandas(idf, preserve_index=False)
    elif isinstance(idf, pa.Table):
        tb = idf
    elif isinstance(idf, ColumnFrame) or getattr(idf, "is_partitioned", False):
        tb = (idf.materialize() if getattr(idf, "is_partitioned", False) else idf).to_arrow()
    else:
        raise TypeError("write_dataset: pass a ColumnFrame, ResultFrame, pandas DataFrame or pyarrow Table")
    if column_order:
        if len(column_order) != len(tb.column_names):
            raise ValueError("Count of column(s) specified in column_order argument do not match Dataframe")
        diff = [c for c in column_order if c not in set(tb.column_names)]
        if diff:
            raise ValueError("Column(s) specified in column_order argument not found in Dataframe: " + str(diff))
        tb = tb.select(list(column_order))
    mode = file_configs.get("mode", "error")
    part = 0
    if os.path.exists(file_path):
        if mode == "error":
            raise FileExistsError(file_path)
        if mode == "overwrite":
            shutil.rmtree(file_path)          # part files AND Spark-style partition sub-directories
        elif mode == "append":
            part = sum(1 for f in os.listdir(file_path) if f.startswith("part-"))
        else:
            raise ValueError("mode must be error, overwrite or append")
    os.makedirs(file_path, exist_ok=True)
    stem = os.path.join(file_path, "part-%05d" % part)
    if file_type == "csv":
        import pyarrow.csv as pacsv           # Arrow keeps nullable integers integers (pandas would print 39.0)
        pacsv.write_csv(tb, stem + ".csv", pacsv.WriteOptions(include_header=_truthy(file_configs.get("header", "false")),
                                                              delimiter=file_configs.get("delimiter", ","), quoting_style="needed"))
    elif file_type == "parquet":
        pq.write_table(tb, stem + ".parquet")
    elif file_type == "json":
        import json
        with open(stem + ".json", "w") as fh:      # JSON Lines; null fields are omitted, like Spark's writer
            for row in tb.to_pylist():
                fh.write(json.dumps({k: v for k, v in row.items() if v is not None}, ensure_ascii=False) + "\n")
    elif file_type == "avro":
        from .avro_io import write_avro
        codec = {"uncompressed": "null", "none": "null"}.get(str(file_configs.get("compression", "snappy")).lower(),
                                                             str(file_configs.get("compression", "snappy")).lower())
        write_avro(tb, stem + ".avro", codec)
    else:
        raise NotImplementedError("file_type %r" % file_type)
    open(os.path.join(file_path, "_SUCCESS"), "w").close()
