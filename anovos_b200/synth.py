"""Synthetic wide tabular frames of SURVEY.md 8(d): generated ON THE DEVICE by the Philox
kernel of libanovos_b200 (anv_synth_f32 / anv_synth_codes), plus a NumPy twin drawing from
the same distribution families for the CPU baseline (not bit-identical: the baseline only
needs the same workload shape)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .frame import ColumnFrame

NULL_RATES = (0.0, 0.001, 0.02, 0.3)
CARDS = (2, 12, 100, 10000)


def column_params(c: int, seed: int, shifted: bool = False):
    """(family, a, b, null_rate) of numeric column c.  Drift target (`shifted`): 1/3 of the
    columns unchanged, 1/3 mean-shifted by 0.25 sigma, 1/3 scale x1.5."""
    rng = np.random.default_rng([seed, c])
    fam = c % 4
    mu, sigma = float(rng.uniform(-50, 50)), float(rng.uniform(0.5, 20))
    lo = float(rng.uniform(-100, 0))
    hi = lo + float(rng.uniform(1, 200))
    if fam == 0:
        a, b = mu, sigma
    elif fam == 1:
        a, b = 0.0, 0.75
    elif fam == 2:
        a, b = lo, hi
    else:
        a, b = 0.0, float(rng.uniform(0.5, 5))
    if shifted:
        mode = c % 3
        if mode == 1:
            if fam == 0:
                a += 0.25 * b
            elif fam == 1:
                a += 0.25 * 0.75
            elif fam == 2:
                a, b = a + 0.07 * (b - a), b + 0.07 * (b - a)
            else:
                b *= 1.2
        elif mode == 2:
            if fam in (0, 1, 3):
                b *= 1.5
            else:
                b = a + 1.5 * (b - a)
    return fam, a, b, NULL_RATES[c % 4]


def device_frame(rows: int, cols: int, seed: int = 42, first_col: int = 0, shifted: bool = False, cat_every: int = 0,
                 prefix: str = "c", row0: int = 0) -> ColumnFrame:
    """`cols` columns starting at global column id `first_col`; every `cat_every`-th column
    (0 = none) is a dictionary-coded string column (Zipf s=1.2).  row0 (multiple of 32): generate
    the row chunk [row0, row0 + rows) of a larger frame, bit-identical to those rows of it."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    data = {}
    words = (rows + 31) // 32
    for i in range(cols):
        c = first_col + i
        name = "%s%04d" % (prefix, c)
        if cat_every and c % cat_every == cat_every - 1:
            card = CARDS[(c // cat_every) % 4]
            rate = NULL_RATES[c % 4]
            x = torch.empty(rows, dtype=torch.int32, device="cuda")
            v = torch.zeros(words, dtype=torch.int32, device="cuda") if rate > 0 else None
            _lib.check(L.anv_synth_codes_rows(x.data_ptr(), v.data_ptr() if v is not None else None, rows, row0, seed, c,
                                              card, 1.2, rate, st), "anv_synth_codes_rows")
            data[name] = (x, v, ["cat_%05d" % k for k in range(card)])
            continue
        fam, a, b, rate = column_params(c, 42, shifted)
        x = torch.empty(rows, dtype=torch.float32, device="cuda")
        v = torch.zeros(words, dtype=torch.int32, device="cuda") if rate > 0 else None
        _lib.check(L.anv_synth_f32_rows(x.data_ptr(), v.data_ptr() if v is not None else None, rows, row0, seed, c, fam,
                                        a, b, rate, st), "anv_synth_f32_rows")
        data[name] = (x, v) if v is not None else x
    return ColumnFrame.from_tensors(data, n_rows=rows)


def host_column(rows: int, c: int, seed: int = 42, shifted: bool = False):
    """NumPy twin of one numeric column -> (float32 values, bool valid)."""
    fam, a, b, rate = column_params(c, 42, shifted)
    rng = np.random.default_rng([seed, c, 7])
    if fam == 0:
        x = rng.normal(a, b, rows)
    elif fam == 1:
        x = np.exp(rng.normal(a, b, rows))
    elif fam == 2:
        x = rng.uniform(a, b, rows)
    else:
        x = np.where(rng.random(rows) < 179 / 256, 0.0, rng.exponential(b, rows))
    valid = rng.random(rows) >= rate if rate > 0 else np.ones(rows, bool)
    return x.astype(np.float32), valid


def host_table(rows: int, cols: int, seed: int = 42, first_col: int = 0, shifted: bool = False, prefix: str = "c"):
    import pyarrow as pa
    arrays, names = [], []
    for i in range(cols):
        c = first_col + i
        x, valid = host_column(rows, c, seed, shifted)
        arrays.append(pa.array(x, mask=None if valid.all() else ~valid))
        names.append("%s%04d" % (prefix, c))
    return pa.table(arrays, names=names)
