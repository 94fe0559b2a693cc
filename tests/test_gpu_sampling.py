"""The device sampler (csrc/sample.cu: Spark's XORShiftRandom stream generated in parallel by GF(2) jump-ahead) against
the oracle's sequential restatement, `data_sample` through the public API, and `drift_detector.statistics` with the
REFERENCE DEFAULTS (use_sampling=True, sample_size=100000, sample_seed=42: drift_detector.py:28-33,187-211) against an
oracle that samples identically."""
import warnings

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

from oracle import api as O
from oracle import spark_semantics as S
from test_data_sampling_cpu import _mixed


@pytest.mark.parametrize("n,seed", [(1, 0), (31, 12), (1024, 42), (1025, 43), (300_001, 2 ** 40 + 5), (70_000, -3)])
def test_sample_mask_equals_sequential_stream(n, seed):
    import torch
    from anovos_b200.data_ingest.data_sampling import fraction_threshold, sample_mask
    k = S.xorshift_uniform53(seed, n)
    thr = fraction_threshold(0.37)
    assert np.array_equal(sample_mask(n, seed, [thr]).cpu().numpy(), k < np.uint64(thr))
    rng = np.random.default_rng(n)
    strata = rng.integers(-1, 4, n).astype(np.int32)            # -1: unknown stratum -> never kept
    thrs = [fraction_threshold(f) for f in (0.1, 1.0, 0.5)]     # stratum 3 is out of range as well
    got = sample_mask(n, seed, thrs, torch.from_numpy(strata).cuda()).cpu().numpy()
    ok = (strata >= 0) & (strata < 3)
    exp = ok & (k < np.asarray(thrs, np.uint64)[np.where(ok, strata, 0)])
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("kw", [dict(method_type="random", fraction=0.3, seed_value=5),
                                dict(method_type="stratified", strata_cols="g|k", fraction=0.4, seed_value=3),
                                dict(method_type="stratified", strata_cols=["g", "k"], stratified_type="balanced", fraction=0.9)])
@pytest.mark.parametrize("parts", [None, [60_000, 1, 89_999]])
def test_data_sample_equals_oracle(kw, parts):
    from anovos.data_ingest.data_sampling import data_sample
    t = _mixed(150_000, 2, parts)
    exp = O.data_sample(t, **kw)
    got = data_sample(t, **kw)
    fr = got.materialize(["id", "x"]) if getattr(got, "is_partitioned", False) else got
    assert np.array_equal(fr.column("id").device()[0].cpu().numpy(), np.asarray(exp.column("id")))
    assert fr.count() == exp.num_rows


def test_reference_dataset_ranges():
    """test_data_sampling.py:20-82 of the reference through the product (35-row dataset restated in the CPU test)."""
    from anovos.data_ingest.data_sampling import data_sample
    from test_data_sampling_cpu import _sample_table
    t = _sample_table()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o3 = data_sample(t, strata_cols="all", method_type="random", fraction=0.5, seed_value=1)
        o4 = data_sample(t, strata_cols="all", method_type="stratified", stratified_type="balanced", fraction=0.75)
    assert 12 < o3.count() < 24
    g = o4.column("gender")
    codes = g.device()[0].cpu().numpy()
    f = int((codes == g.dictionary.index("F")).sum())
    assert 6 < f < 13 and 6 < o4.count() - f < 13


@pytest.mark.parametrize("sample_method,extra", [("random", {}), ("stratified", {"strata_cols": "g|k"})])
def test_statistics_with_reference_defaults_equals_oracle(tmp_path, sample_method, extra):
    import anovos.drift_stability.drift_detector as dd
    src, tgt = _mixed(400_003, 7).drop_columns(["id"]), _mixed(250_001, 8).drop_columns(["id"])
    kw = dict(method_type="all", sample_method=sample_method, **extra)        # use_sampling=True, sample_size=100000, seed 42
    got = dd.statistics(None, tgt, src, source_path=str(tmp_path / "g"), **kw).toPandas()
    exp = O.statistics(tgt, src, source_path=str(tmp_path / "o"), **kw)
    assert got["attribute"].tolist() == exp["attribute"].tolist()
    for m in ("PSI", "HD", "JSD", "KS"):
        assert np.allclose(got[m].values.astype(float), np.asarray(exp[m], float), rtol=1e-9, atol=1e-12), m
    assert got["flagged"].tolist() == exp["flagged"].tolist()
