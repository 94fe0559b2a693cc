"""Synthetic wide tabular frames of SURVEY.md 8(d): generated ON THE DEVICE by the Philox
kernel of libanovos_b200 (anv_synth_f32 / anv_synth_codes), plus a BIT-IDENTICAL NumPy twin (host_column / host_codes /
host_table) so that the CPU oracle and the GPU see the same values."""
from __future__ import annotations

import ctypes as C
import functools
import math

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


def numeric_ordinal(c: int, cat_every: int) -> int:
    """Index of numeric column c among the NUMERIC columns of a mixed frame (every cat_every-th column is
    categorical): family and null rate cycle over the numeric columns, so a mixed frame (75 % numeric /
    25 % categorical) still holds all four families and null rates.  The Philox key stays the global column id."""
    return c - c // cat_every if cat_every else c


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
    fam, a, b, rate = column_params(numeric_ordinal(c, cat_every), 42, shifted)

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


# ---- bit-identical NumPy twin of csrc/synth.cu ----------------------------------------------------------
# Philox4x32-10 keyed by (seed, column), counter = (row / 4, stream), and the same correctly rounded float32
# operation sequence as the device code (every np.float32 add / multiply / divide / sqrt is IEEE-754 exact, and the
# kernel uses the _rn intrinsics so nvcc cannot contract them into FMAs): host_column(rows, c) == the device
# column bit for bit.  tests/test_gpu_parity_scale.py checks exactly that, so the CPU oracle, the reference arm of
# bench.py and the GPU all see identical values.

_F = np.float32
_M32 = np.uint64(0xFFFFFFFF)


def _philox_key(seed: int, column: int):
    k = (int(seed) ^ (0x9E3779B97F4A7C15 * (int(column) + 1))) & 0xFFFFFFFFFFFFFFFF
    return k & 0xFFFFFFFF, k >> 32


def _philox4x32_10(g: np.ndarray, stream: int, key):
    """g: uint64 counters (global row group = row // 4) -> four uint32 arrays (x, y, z, w)."""
    c0 = (g & _M32).astype(np.uint64)
    c1 = (g >> np.uint64(32)).astype(np.uint64)
    c2 = np.full(g.shape, stream, np.uint64)
    c3 = np.zeros(g.shape, np.uint64)
    k0, k1 = int(key[0]), int(key[1])
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    s32 = np.uint64(32)
    for _ in range(10):
        p0, p1 = m0 * c0, m1 * c2
        c0, c1, c2, c3 = (p1 >> s32) ^ c1 ^ np.uint64(k0), p1 & _M32, (p0 >> s32) ^ c3 ^ np.uint64(k1), p0 & _M32
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _mad(a, b, c):
    return (a * b).astype(np.float32) + c   # two roundings, like mad() of synth.cu (never an FMA)


def _u01(b):
    return ((b >> np.uint32(8)).astype(np.float32) + _F(0.5)) * _F(5.9604644775390625e-08)


def _det_log(x):
    bits = x.view(np.uint32)
    e = (bits >> np.uint32(23)).astype(np.int32) - np.int32(127)
    m = ((bits & np.uint32(0x007FFFFF)) | np.uint32(0x3F800000)).view(np.float32)
    big = m > _F(1.41421354)
    m = np.where(big, m * _F(0.5), m).astype(np.float32)
    e = e + big.astype(np.int32)
    s = (m - _F(1.0)) / (m + _F(1.0))
    z = s * s
    p = np.full(x.shape, _F(0.111111112), np.float32)
    for c in (0.142857149, 0.2, 0.333333343, 1.0):
        p = _mad(p, z, _F(c))
    lnm = (_F(2.0) * s) * p
    return _mad(e.astype(np.float32), _F(0.693147182), lnm)


def _det_exp(y):
    k = np.floor(_mad(y, _F(1.44269502), _F(0.5)))
    r = y - k * _F(0.693359375)
    r = r - k * _F(-2.12194440e-4)
    p = np.full(y.shape, _F(1.38888892e-3), np.float32)
    for c in (8.33333377e-3, 4.16666679e-2, 0.166666672, 0.5, 1.0, 1.0):
        p = _mad(p, r, _F(c))
    ki = np.clip(k.astype(np.int32), -126, 127)
    scale = ((ki + np.int32(127)).astype(np.uint32) << np.uint32(23)).view(np.float32)
    return p * scale


def _det_sincos2pi(u):
    t = u * _F(4.0)
    q = t.astype(np.int32)
    f = t - q.astype(np.float32)
    th = (f - _F(0.5)) * _F(1.57079637)
    z = th * th
    sp = np.full(u.shape, _F(2.75573188e-6), np.float32)
    for c in (-1.98412701e-4, 8.33333377e-3, -0.166666672, 1.0):
        sp = _mad(sp, z, _F(c))
    sn = th * sp
    cp = np.full(u.shape, _F(-2.75573200e-7), np.float32)
    for c in (2.48015876e-5, -1.38888892e-3, 4.16666679e-2, -0.5):
        cp = _mad(cp, z, _F(c))
    cs = _mad(cp, z, _F(1.0))
    a = (cs - sn) * _F(0.707106769)
    b = (cs + sn) * _F(0.707106769)
    qq = q & 3
    cos = np.where(qq == 0, a, np.where(qq == 1, -b, np.where(qq == 2, -a, b))).astype(np.float32)
    sin = np.where(qq == 0, b, np.where(qq == 1, a, np.where(qq == 2, -b, -a))).astype(np.float32)
    return sin, cos


def _normals4(x, y, z, w):
    r0 = np.sqrt(_F(-2.0) * _det_log(_u01(x)))
    r1 = np.sqrt(_F(-2.0) * _det_log(_u01(z)))
    s0, c0 = _det_sincos2pi(_u01(y))
    s1, c1 = _det_sincos2pi(_u01(w))
    return r0 * c0, r0 * s0, r1 * c1, r1 * s1


def _interleave(parts, rows):
    out = np.empty((parts[0].shape[0], 4), parts[0].dtype)
    for i, p in enumerate(parts):
        out[:, i] = p
    return out.reshape(-1)[:rows]


def _twin_chunk(n_rows, row0, seed, column, kind, a, b, null_rate, card=0, zipf_s=1.2):
    """Rows [row0, row0 + n_rows) (row0 % 4 == 0) of a column -> (values, valid bool | None)."""
    key = _philox_key(seed, column)
    n4 = (n_rows + 3) // 4
    g = np.arange(n4, dtype=np.uint64) + np.uint64(row0 >> 2)
    x, y, z, w = _philox4x32_10(g, 0, key)
    a, b = _F(a), _F(b)
    with np.errstate(all="ignore"):
        if kind == "codes":
            oms = 1.0 - float(_F(zipf_s))
            span = _F(math.pow(card + 1.0, oms) - 1.0)
            inv = _F(1.0 / oms)
            parts = []
            for r in (x, y, z, w):
                v = _det_exp(inv * _det_log(_mad(_u01(r), span, _F(1.0))))
                parts.append(np.minimum(np.maximum(v.astype(np.int32) - 1, 0), card - 1).astype(np.int32))
        elif kind == 0:
            parts = [_mad(v, b, a) for v in _normals4(x, y, z, w)]
        elif kind == 1:
            parts = [_det_exp(_mad(v, b, a)) for v in _normals4(x, y, z, w)]
        elif kind == 2:
            wd = b - a
            parts = [_mad(_u01(r), wd, a) for r in (x, y, z, w)]
        else:
            parts = [np.where((r & np.uint32(0xFF)) < 179, _F(0.0), (-b) * _det_log(_u01(r))).astype(np.float32)
                     for r in (x, y, z, w)]
    vals = _interleave(parts, n_rows)
    valid = None
    if null_rate > 0:
        thr = np.uint32(min(float(_F(null_rate) * _F(4294967296.0)), 4294967040.0))
        nx = _philox4x32_10(g, 1, key)
        valid = _interleave([r >= thr for r in nx], n_rows)
    return vals, valid


def _twin(rows, row0, chunk, **kw):
    if rows == 0:
        z = np.zeros(0, np.int32 if kw["kind"] == "codes" else np.float32)
        return z, (np.zeros(0, bool) if kw["null_rate"] > 0 else None)
    vs, ms = [], []
    for s in range(0, rows, chunk):
        v, m = _twin_chunk(min(chunk, rows - s), row0 + s, **kw)
        vs.append(v)
        ms.append(m)
    return (np.concatenate(vs) if len(vs) > 1 else vs[0]), (None if ms[0] is None else (np.concatenate(ms) if len(ms) > 1 else ms[0]))


def host_column(rows: int, c: int, seed: int = 42, shifted: bool = False, row0: int = 0, cat_every: int = 0):
    """Bit-identical NumPy twin of numeric column c of device_frame(rows, ..., seed, shifted, cat_every) ->
    (float32 values, bool valid).  row0 (multiple of 4): the rows [row0, row0 + rows) of a larger frame."""
    fam, a, b, rate = column_params(numeric_ordinal(c, cat_every), 42, shifted)
    v, m = _twin(rows, row0, 1 << 22, seed=seed, column=c, kind=fam, a=a, b=b, null_rate=rate)
    return v, (np.ones(rows, bool) if m is None else m)


def host_codes(rows: int, c: int, cat_every: int, seed: int = 42, row0: int = 0):
    """Bit-identical twin of the dictionary-coded string column c -> (int32 codes, bool valid, dictionary)."""
    card = CARDS[(c // cat_every) % 4]
    rate = NULL_RATES[c % 4]
    v, m = _twin(rows, row0, 1 << 22, seed=seed, column=c, kind="codes", a=0.0, b=0.0, null_rate=rate, card=card)
    return v, (np.ones(rows, bool) if m is None else m), ["cat_%05d" % k for k in range(card)]


def host_table(rows: int, cols: int, seed: int = 42, first_col: int = 0, shifted: bool = False, prefix: str = "c",
               cat_every: int = 0, columns=None):
    """pyarrow Table holding the same values as device_frame(rows, cols, seed, first_col, shifted, cat_every)
    (`columns`: only these global column ids)."""
    import pyarrow as pa
    arrays, names = [], []
    for c in (columns if columns is not None else range(first_col, first_col + cols)):
        if cat_every and c % cat_every == cat_every - 1:
            codes, valid, dic = host_codes(rows, c, cat_every, seed)
            arr = pa.DictionaryArray.from_arrays(pa.array(codes, mask=None if valid.all() else ~valid), pa.array(dic)).cast(pa.string())
        else:
            x, valid = host_column(rows, c, seed, shifted, cat_every=cat_every)
            arr = pa.array(x, mask=None if valid.all() else ~valid)
        arrays.append(arr)
        names.append("%s%04d" % (prefix, c))
    return pa.table(arrays, names=names)
