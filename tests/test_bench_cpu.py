"""bench.py pieces that run without a GPU: the `parity` leg (the step's result frames vs the oracle on the bit-identical NumPy
twin) and the N > 1 summary matrix on the frames of a MIXED workload - exercised through the NumPy engine stand-in."""
import tempfile

import numpy as np

import cpu_engine


def test_parity_leg_and_summary_matrix_on_a_mixed_frame():
    import bench
    import anovos.drift_stability.drift_detector as dd
    from anovos_b200 import parallel, synth
    from anovos_b200.frame import ColumnFrame
    rows, cols, cat_every = 20_000, 8, 4
    with cpu_engine.installed():
        src = ColumnFrame.from_arrow(synth.host_table(rows, cols, cat_every=cat_every))
        tgt = ColumnFrame.from_arrow(synth.host_table(rows, cols, seed=43, shifted=True, cat_every=cat_every))
        frames = bench.stats_step(src)
        drift = dd.statistics(None, tgt, src, method_type="all", use_sampling=False, source_path=tempfile.mkdtemp()).toPandas()
        par = bench.parity_check(rows, cols, 0, cat_every, src, frames, drift)
    assert par["mismatches"] == 0 and par["cells_checked"] > 100, par
    assert set(par["columns"]) == {"c0000", "c0001", "c0002", "c0003", "c0004"}       # four numeric families + one string column
    assert par["drift"]["max_rel_err"] < 1e-9
    # the per-rank summary matrix of the N > 1 exchange: frames of different row counts align on `attribute`
    m, names = parallel.frames_to_matrix(frames)
    assert m.shape == (cols, len(names)) and "stddev" in names and "fill_count" in names
    string_rows = [i for i, c in enumerate(frames[0]["attribute"]) if c in ("c0003", "c0007")]
    assert np.isnan(m[string_rows][:, names.index("stddev")]).all()
