"""Synthetic wide tabular frames of SURVEY.md 8(d): generated ON THE DEVICE by the Philox
kernel of libanovos_b200 (anv_synth_f32 / anv_synth_codes), plus a NumPy twin drawing from
the same distribution families for the CPU baseline (not bit-identical: the baseline only
needs the same workload shape)."""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np

from . import _lib
from .frame import ColumnFrame

NULL_RATES = (0.0, 0.001, 0.02, 0.3)
CARDS = (2, 12, 100, 10000)


@functools.lru_cache(maxsize=None)
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


_DICTS = {}  # cardinality -> dictionary strings (shared by every chunk: built once)


def _column_generator(rows: int, c: int, seed: int, shifted: bool, cat_every: int, row0: int):
    """-> (spark dtype, dictionary | None, has_nulls, loader) of global column id c; loader() launches the
    Philox kernel on the current stream and returns (values, validity words | None) on the device."""
    L = _lib.lib()
    words = (rows + 31) // 32
    if cat_every and c % cat_every == cat_every - 1:
        card = CARDS[(c // cat_every) % 4]
        rate = NULL_RATES[c % 4]

        def load_codes():
            torch = _lib.require_cuda()
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            x = torch.empty(rows, dtype=torch.int32, device="cuda")
            v = torch.zeros(words, dtype=torch.int32, device="cuda") if rate > 0 else None
            _lib.check(L.anv_synth_codes_rows(x.data_ptr(), v.data_ptr() if v is not None else None, rows, row0, seed, c,
                                              card, 1.2, rate, st), "anv_synth_codes_rows")
            return x, v
        if card not in _DICTS:
            _DICTS[card] = ["cat_%05d" % k for k in range(card)]
        return "string", _DICTS[card], rate > 0, load_codes
    fam, a, b, rate = column_params(c, 42, shifted)

    def load_f32():
        torch = _lib.require_cuda()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        x = torch.empty(rows, dtype=torch.float32, device="cuda")
        v = torch.zeros(words, dtype=torch.int32, device="cuda") if rate > 0 else None
        _lib.check(L.anv_synth_f32_rows(x.data_ptr(), v.data_ptr() if v is not None else None, rows, row0, seed, c, fam,
                                        a, b, rate, st), "anv_synth_f32_rows")
        return x, v
    return "float", None, rate > 0, load_f32


def device_frame(rows: int, cols: int, seed: int = 42, first_col: int = 0, shifted: bool = False, cat_every: int = 0,
                 prefix: str = "c", row0: int = 0, lazy: bool = False) -> ColumnFrame:
    """`cols` columns starting at global column id `first_col`; every `cat_every`-th column
    (0 = none) is a dictionary-coded string column (Zipf s=1.2).  row0 (multiple of 32): generate
    the row chunk [row0, row0 + rows) of a larger frame, bit-identical to those rows of it.
    lazy: columns are generated on first use and can be dropped again (streamed chunks)."""
    from collections import OrderedDict
    from .frame import Column
    _lib.require_cuda()
    out = OrderedDict()
    for i in range(cols):
        c = first_col + i
        name = "%s%04d" % (prefix, c)
        sd, dic, nulls, loader = _column_generator(rows, c, seed, shifted, cat_every, row0)
        col = Column(name, sd, rows, anv_dtype=_lib.ANV_I32 if dic is not None else _lib.ANV_F32, dictionary=dic,
                     null_count=None if nulls else 0, loader=loader)
        if not lazy:
            col.device()
            col._loader = None
        out[name] = col
    return ColumnFrame(out, rows)


def partitioned_frame(rows: int, cols: int, chunk_rows: int, seed: int = 42, first_col: int = 0, shifted: bool = False,
                      cat_every: int = 0, group=None):
    """The same frame as device_frame(rows, ...), never resident: a PartitionedFrame whose chunks are
    (re)generated column by column when a pass touches them and freed afterwards (C4/C5 streaming)."""
    from .partitioned import PartitionedFrame
    chunk_rows = max(32, int(chunk_rows) // 32 * 32)
    starts = list(range(0, rows, chunk_rows)) or [0]
    schema = device_frame(0, cols, seed, first_col, shifted, cat_every, lazy=True)
    return PartitionedFrame(schema, [min(chunk_rows, rows - s) for s in starts],
                            lambda i: device_frame(min(chunk_rows, rows - starts[i]), cols, seed, first_col, shifted,
                                                   cat_every, row0=starts[i], lazy=True), group=group)


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
