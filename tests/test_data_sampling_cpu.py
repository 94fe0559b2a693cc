"""data_sample (reference data_ingest/data_sampling.py:8-149) under `-m "not gpu"`: the oracle restatement of Spark's
XORShiftRandom / Bernoulli / sampleBy samplers against the reference's own test (count ranges on its 35-row dataset,
/root/reference/src/test/anovos/data_ingest/test_data_sampling.py:20-82; the rows are restated below), the host-side
pieces of the product (seed hashing in the C library, exact integer thresholds), and the product's host path through
the engine stand-in (== oracle row for row).  `-m gpu`: tests/test_gpu_sampling.py runs the device kernel."""
import warnings

import numpy as np
import pyarrow as pa
import pytest

import cpu_engine
from oracle import api as O
from oracle import spark_semantics as S

# data/data_sample/test_data_sample.csv of the reference: gender, label_1, label_2, label_3 (35 rows)
_ROWS = [("F", "T", "T", "T")] * 8 + [("F", "F", "F", "F")] * 8 + [("F", "T", "F", "T")] * 8 + \
        [("M", "T", "T", "T")] * 4 + [("M", "F", "F", "F")] * 4 + [("M", "T", "F", "T")] * 4


def _sample_table():
    return pa.table({n: [r[i] for r in _ROWS] for i, n in enumerate(["gender", "label_1", "label_2", "label_3"])})


def _count(t, g):
    return sum(1 for v in t.column("gender").to_pylist() if v == g)


def test_reference_data_sampling_test_through_the_oracle():
    t = _sample_table()
    o1 = O.data_sample(t, strata_cols="all", method_type="stratified", fraction=0.75)
    assert o1.column_names == ["gender", "label_1", "label_2", "label_3"]
    assert 12 < _count(o1, "F") < 30 and 3 < _count(o1, "M") < 15
    o2 = O.data_sample(t, strata_cols="all", method_type="stratified", fraction=0.5, seed_value=1)
    assert 8 < _count(o2, "F") < 20 and 2 < _count(o2, "M") < 10
    o3 = O.data_sample(t, strata_cols="all", method_type="random", fraction=0.5, seed_value=1)
    assert 12 < o3.num_rows < 24
    o4 = O.data_sample(t, strata_cols="all", method_type="stratified", stratified_type="balanced", fraction=0.75)
    assert 6 < _count(o4, "F") < 13 and 6 < _count(o4, "M") < 13


def test_argument_checks_match_the_reference():
    t = _sample_table()
    for kw in ({"fraction": "0.1"}, {"fraction": 0}, {"fraction": 1.5}, {"seed_value": 1.0}, {"method_type": "x"},
               {"method_type": "stratified", "unique_threshold": "a"}, {"method_type": "stratified", "unique_threshold": 1.5},
               {"method_type": "stratified", "unique_threshold": 0}, {"method_type": "stratified", "stratified_type": "x"},
               {"method_type": "stratified", "strata_cols": "nope"}, {"method_type": "stratified", "strata_cols": "gender", "drop_cols": "gender"}):
        with pytest.raises(TypeError):
            O.data_sample(t, **kw)
        from anovos.data_ingest.data_sampling import data_sample
        with cpu_engine.installed(), pytest.raises(TypeError):
            data_sample(t, **kw)


def test_generator_pieces():
    from anovos_b200 import _lib
    from anovos_b200.data_ingest.data_sampling import fraction_threshold
    L = _lib.lib()
    for seed in (0, 1, 12, 42, -7, 2 ** 40 + 3, -2 ** 63, 2 ** 63 - 1):
        assert L.anv_spark_hash_seed(seed) == S.xorshift_hash_seed(seed)
    k = S.xorshift_uniform53(42, 2000)
    assert k.max() < 2 ** 53 and 0.45 < (k.astype(np.float64) * 2.0 ** -53).mean() < 0.55
    for f in (0.1, 0.5, 1.0, 1e-9, 100000 / 400003, 0.3333333333333333):
        thr = fraction_threshold(f)
        x = k.astype(np.float64) * 2.0 ** -53
        assert np.array_equal(x < f, k < np.uint64(thr))                   # integer compare == Spark's double compare
        for kk in (thr - 1, thr):
            if 0 <= kk < 2 ** 53:
                assert (kk * 2.0 ** -53 < f) == (kk < thr)


def _mixed(n, seed, parts=None):
    rng = np.random.default_rng(seed)
    t = pa.table({"x": pa.array(rng.normal(0, 1, n).astype(np.float32), mask=rng.random(n) < 0.05),
                  "g": pa.array(rng.choice(["a", "b", "cc"], n, p=[0.6, 0.3, 0.1]), mask=rng.random(n) < 0.02),
                  "k": pa.array(rng.integers(0, 3, n).astype(np.int32)),
                  "id": pa.array(np.arange(n, dtype=np.int64))})
    return O.with_spark_partitions(t, parts) if parts else t


@pytest.mark.parametrize("kw", [dict(method_type="random", fraction=0.3, seed_value=5),
                                dict(method_type="stratified", strata_cols="g|k", fraction=0.4, seed_value=3),
                                dict(method_type="stratified", strata_cols="all", fraction=0.5),
                                dict(method_type="stratified", strata_cols=["g", "k"], stratified_type="balanced", fraction=0.9)])
@pytest.mark.parametrize("parts", [None, [1500, 1, 2499]])
def test_product_host_path_equals_oracle(kw, parts):
    from anovos.data_ingest.data_sampling import data_sample
    t = _mixed(4000, 1, parts)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")       # "all": id and x are dropped from the strata (high cardinality)
        exp = O.data_sample(t, **kw)
        with cpu_engine.installed():
            got = data_sample(t, **kw)
            ids = got.materialize(["id"]).column("id").device()[0].numpy() if getattr(got, "is_partitioned", False) \
                else got.column("id").device()[0].numpy()
    assert np.array_equal(ids, np.asarray(exp.column("id")))
    assert 0 < len(ids) < 4000
