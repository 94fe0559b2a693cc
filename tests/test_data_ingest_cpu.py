"""N4 (formats either side of the path): csv / parquet -> ColumnFrame typing rules and the writer layout.
CPU only: frames are built but no kernel is launched."""
import os

import pyarrow as pa
import pyarrow.parquet as pq
import pytest


def test_read_csv_and_parquet_typing(tmp_path, income):
    from anovos.data_ingest.data_ingest import read_dataset, write_dataset
    import pyarrow.csv as pacsv
    small = income.slice(0, 500)
    d = tmp_path / "csvdir"
    os.makedirs(d)
    pacsv.write_csv(small, str(d / "part-00000.csv"))
    fr = read_dataset(None, str(d), "csv", {"header": "True", "delimiter": ",", "inferSchema": "True"})
    types = dict(fr.dtypes)
    assert fr.count() == 500 and types["age"] == "int" and types["logfnl"] == "double" and types["workclass"] == "string"
    from anovos.shared.utils import attributeType_segregation
    num, cat, other = attributeType_segregation(fr)
    assert "age" in num and "workclass" in cat and not other
    raw = read_dataset(None, str(d), "csv", {"header": "True"})          # no inferSchema: all strings, like Spark
    assert set(t for _, t in raw.dtypes) == {"string"}
    p = tmp_path / "pq"
    os.makedirs(p)
    pq.write_table(small, str(p / "part-00000.snappy.parquet"))
    fp = read_dataset(None, str(p), "parquet")
    assert fp.columns == small.column_names and dict(fp.dtypes)["fnlwgt"] == "int"
    with pytest.raises(FileNotFoundError):
        read_dataset(None, str(p), "avro")            # a directory without .avro part files
    with pytest.raises(NotImplementedError):
        read_dataset(None, str(p), "orc")
    # writer: Spark-style directory with a part file and _SUCCESS, mode error / overwrite
    import pandas as pd
    out = tmp_path / "out"
    df = pd.DataFrame({"attribute": ["a", "b"], "mean": [1.5, 2.5]})
    write_dataset(df, str(out), "csv", {"header": "True", "mode": "overwrite"})
    assert sorted(os.listdir(out)) == ["_SUCCESS", "part-00000.csv"]
    back = read_dataset(None, str(out), "csv", {"header": "True", "inferSchema": "True"})
    assert back.columns == ["attribute", "mean"] and back.count() == 2
    with pytest.raises(FileExistsError):
        write_dataset(df, str(out), "csv", {"header": "True"})
    with pytest.raises(ValueError):
        write_dataset(df, str(tmp_path / "o2"), "csv", column_order=["mean"])


def test_spark_scan_partitions(tmp_path):
    """FilePartition arithmetic (maxSplitBytes, size-descending files, next-fit packing) and Hadoop's line-split rule:
    the income CSV of the reference (5 891 566 bytes) splits at 4 MiB for local[>=3] -> the 22 984 + 9 577 rows recorded
    in tests/golden/income_partitions.json."""
    import json
    import numpy as np
    from conftest import GOLDEN
    from anovos_b200.data_ingest import data_ingest as di
    gold = json.load(open(os.path.join(GOLDEN, "income_partitions.json")))
    big = tmp_path / "big.csv"
    rng = np.random.default_rng(0)
    lines = ["a,b,c"] + ["%d,%s,%.6f" % (i, "x" * int(rng.integers(0, 40)), rng.random()) for i in range(260_000)]
    big.write_text("\n".join(lines) + "\n")
    size = os.path.getsize(big)
    assert size > 2 * di.OPEN_COST
    parts = di.spark_file_partitions([str(big)], 16)
    assert [(s, l) for p in parts for _, s, l in p] == [(o, min(di.OPEN_COST, size - o)) for o in range(0, size, di.OPEN_COST)]
    rows = [sum(di.csv_rows_in_range(f, s, l, True) for f, s, l in p) for p in parts]
    assert sum(rows) == 260_000 and len(rows) == -(-size // di.OPEN_COST)
    starts = np.cumsum([0] + [len(x) + 1 for x in lines])[1:-1]            # data-line start offsets
    brute = [int(((starts <= e) & (starts > (e - di.OPEN_COST if e > di.OPEN_COST else -1))).sum())
             for e in range(di.OPEN_COST, size + di.OPEN_COST, di.OPEN_COST)]
    assert rows == brute
    assert len(di.spark_file_partitions([str(big)], 1)) == 1                  # one core: one split of the whole file
    # small files are packed one per partition once their open cost is counted, largest first
    small = []
    for i, n in enumerate((3000, 5000, 4000)):
        p = tmp_path / ("s%d.csv" % i)
        p.write_text("a\n" + "1\n" * n)
        small.append(str(p))
    packed = di.spark_file_partitions(small, 8)
    assert [os.path.basename(p[0][0]) for p in packed] == ["s1.csv", "s2.csv", "s0.csv"] and all(len(p) == 1 for p in packed)
    fr = di.read_dataset(None, str(big), "csv", {"header": "True", "inferSchema": "True", "spark_cores": 16})
    assert fr.spark_partitions and fr.chunk_rows == rows and fr.count() == 260_000 and fr.columns == ["a", "b", "c"]
    assert gold["split_bytes"] == di.OPEN_COST and gold["file_bytes"] == 5891566 and gold["rows_per_partition"] == [22984, 9577]


def test_flatten_and_transpose_dataframe():
    """shared/utils.py:6-45 on a summary()-shaped frame: explode(create_map) row order, pivot columns sorted."""
    import pandas as pd
    from anovos.shared.utils import flatten_dataframe, transpose_dataframe
    df = pd.DataFrame({"summary": ["count", "mean", "max"], "age": [4, 42.75, 55], "income": [3, 7333.3, 9000]})
    flat = flatten_dataframe(df, ["summary"]).toPandas()
    assert flat.columns.tolist() == ["summary", "key", "value"]
    assert flat[["summary", "key"]].values.tolist() == [["count", "age"], ["count", "income"], ["mean", "age"], ["mean", "income"],
                                                          ["max", "age"], ["max", "income"]]
    tr = transpose_dataframe(df, "summary").toPandas()
    assert tr.columns.tolist() == ["key", "count", "max", "mean"]
    assert tr.values.tolist() == [["age", 4.0, 55.0, 42.75], ["income", 3.0, 9000.0, 7333.3]]


def test_json_avro_round_trip_and_frame_writer(tmp_path, income):
    """file types "json" and "avro" of read_dataset / write_dataset (reference data_ingest.py:23-110), ColumnFrame ->
    file (ADVICE: frames returned by attribute_binning / outlier treatment must be writable), overwrite of a directory
    that holds sub-directories, append."""
    import numpy as np
    from anovos.data_ingest.data_ingest import read_dataset, write_dataset
    from anovos_b200.frame import ColumnFrame
    small = income.slice(0, 300).select(["age", "workclass", "logfnl", "fnlwgt", "income"])
    for ftype, cfg in (("json", {}), ("avro", {}), ("avro", {"compression": "deflate"}), ("avro", {"compression": "uncompressed"}),
                       ("parquet", {}), ("csv", {"header": "True"})):
        out = tmp_path / (ftype + "_" + cfg.get("compression", "x"))
        write_dataset(small, str(out), ftype, dict(cfg, mode="overwrite"))
        fr = read_dataset(None, str(out), ftype, {"header": "True", "inferSchema": "True"})
        assert fr.count() == 300 and sorted(fr.columns) == sorted(small.column_names)
        if ftype == "json":
            assert fr.columns == sorted(small.column_names)          # Spark's JSON inference sorts the fields by name
            assert dict(fr.dtypes)["age"] == "bigint"                # ... and reads integers as LongType
        else:
            assert dict(fr.dtypes)["age"] == "int"
        back = fr.to_arrow()
        for c in small.column_names:
            a, b = back.column(c).to_pylist(), small.column(c).to_pylist()
            assert all((x is None and y is None) or x == y or (isinstance(y, float) and abs(x - y) <= 1e-12 * abs(y))
                       for x, y in zip(a, b)), (ftype, c)
    # a ColumnFrame goes out the same way (no device needed for a host-resident frame)
    fr = ColumnFrame.from_arrow(small)
    d = tmp_path / "frame_out"
    os.makedirs(d / "nested=1")
    open(d / "nested=1" / "part-00000.csv", "w").close()
    write_dataset(fr, str(d), "parquet", {"mode": "overwrite"})       # rmtree: sub-directories do not break overwrite
    assert sorted(os.listdir(d)) == ["_SUCCESS", "part-00000.parquet"]
    write_dataset(fr, str(d), "parquet", {"mode": "append"})
    assert read_dataset(None, str(d), "parquet").count() == 600
    assert pq.read_table(str(d / "part-00000.parquet")).equals(small)


def test_save_stats_layout(tmp_path):
    """report_preprocessing.save_stats (:40-128): <master_path>/<function_name>.csv, header, no index; reread."""
    import pandas as pd
    from anovos.data_report.report_preprocessing import save_stats
    from anovos_b200.result import ResultFrame
    df = pd.DataFrame({"attribute": ["age", "fnlwgt"], "mean": [38.5816, 189778.3665], "mode": ["36", None]})
    save_stats(None, ResultFrame(df), str(tmp_path / "stats"), "measures_of_centralTendency")
    f = tmp_path / "stats" / "measures_of_centralTendency.csv"
    assert f.read_text().splitlines()[0] == "attribute,mean,mode"
    assert pd.read_csv(f).equals(pd.read_csv(f)) and len(pd.read_csv(f)) == 2
    back = save_stats(None, df, str(tmp_path / "stats"), "again", reread=True)
    assert back.columns == ["attribute", "mean", "mode"] and back.count() == 2
    with pytest.raises(NotImplementedError):
        save_stats(None, df, str(tmp_path), "x", run_type="emr")


def test_string_columns_keep_narrow_codes_on_the_host():
    """from_arrow stores dictionary codes in uint8 / int16 / int32 by cardinality (fewer bytes over PCIe; the upload widens
    them): the statistics, the arrow round trip and row slices are those of int32 codes."""
    import numpy as np
    import pyarrow as pa
    import cpu_engine
    from golden_util import assert_frames_match
    from oracle import api as O
    from anovos_b200.frame import ColumnFrame, narrow_code_dtype
    import anovos.data_analyzer.stats_generator as sg
    assert [np.dtype(narrow_code_dtype(k)).itemsize for k in (0, 1, 256, 257, 32768, 32769, 10 ** 6)] == [1, 1, 1, 2, 2, 4, 4]
    n = 70_001
    rng = np.random.default_rng(2)
    cards = [3, 256, 300, 36_000]
    t = pa.table({**{"s%d" % k: pa.array(["v%05d" % v for v in rng.permutation(np.arange(n) % k)], mask=rng.random(n) < 0.1) for k in cards},
                  "x": pa.array(rng.normal(0, 1, n))})
    fr = ColumnFrame.from_arrow(t)
    assert [fr.column("s%d" % k)._host.dtype.itemsize for k in cards] == [1, 1, 2, 4]
    assert fr.to_arrow().equals(t)
    assert fr.slice_rows(64, 5000).to_arrow().equals(t.slice(64, 5000 - 64))
    with cpu_engine.installed():
        for f in ("measures_of_counts", "measures_of_centralTendency", "measures_of_cardinality"):
            assert_frames_match(getattr(sg, f)(None, fr).toPandas(), getattr(O, f)(t))
        wide = ColumnFrame.from_tensors({k: (fr.column(k)._host.astype(np.int32), fr.column(k)._host_valid, fr.column(k).dictionary)
                                         for k in fr.columns if k != "x"}, n_rows=n)
        narrow = ColumnFrame.from_tensors({k: (fr.column(k)._host, fr.column(k)._host_valid, fr.column(k).dictionary)
                                           for k in fr.columns if k != "x"}, n_rows=n)
        assert sg.measures_of_centralTendency(None, wide).toPandas().equals(sg.measures_of_centralTendency(None, narrow).toPandas())
