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
    with pytest.raises(NotImplementedError):
        read_dataset(None, str(p), "avro")
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
