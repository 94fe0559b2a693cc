"""Host driver of the CUDA kernels: turns ColumnFrames + binning models into C-ABI calls.

All device work goes through libanovos_b200.so (include/anovos_b200.h) on torch's
current CUDA stream.  torch is plumbing only: device buffers and streams.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .frame import ColumnFrame

_NP_OF_ANV = {_lib.ANV_F32: np.float32, _lib.ANV_F64: np.float64, _lib.ANV_I32: np.int32, _lib.ANV_I64: np.int64}

MOMENT_FIELDS = ("n_valid", "n_nonzero", "min", "max", "mean", "m2", "m3", "m4")
_MOM_DT = np.dtype([("n_valid", "<i8"), ("n_nonzero", "<i8"), ("min", "<f8"), ("max", "<f8"),
                    ("mean", "<f8"), ("m2", "<f8"), ("m3", "<f8"), ("m4", "<f8")])
_DRIFT_DT = np.dtype([("psi", "<f8"), ("hd", "<f8"), ("jsd", "<f8"), ("ks", "<f8"), ("n_rows", "<i4"), ("r", "<i4")])
_SPEC_DT = np.dtype([("n_bins", "<i4"), ("mode", "<i4"), ("lo", "<f8"), ("inv_w", "<f8"), ("cut_offset", "<i8")])

launch_count = 0  # kernels launched through this module (bench.py reports it)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_bytes(nbytes):
    import torch
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device="cuda")


def _to_dev(arr: np.ndarray):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).cuda()


# ---- K1 ---------------------------------------------------------------------------------

def moments(frame: ColumnFrame, names):
    """-> structured ndarray (one row per name) with MOMENT_FIELDS.  One fused pass."""
    global launch_count
    torch = _lib.require_cuda()
    L = _lib.lib()
    names = list(names)
    if not names:
        return np.zeros(0, dtype=_MOM_DT)
    desc, keep = frame.descriptors(names)
    ws_bytes = L.anv_moments_workspace_bytes(len(names), frame.n_rows)
    ws = _dev_bytes(ws_bytes)
    out = _dev_bytes(len(names) * _MOM_DT.itemsize)
    _lib.check(L.anv_moments(desc.data_ptr(), len(names), frame.n_rows, out.data_ptr(), ws.data_ptr(), ws_bytes,
                             _stream()), "anv_moments")
    launch_count += 2
    return out.cpu().numpy().view(_MOM_DT).copy()


# ---- binning model -> device specs ---------------------------------------------------------

def _round_down_f32(c: float) -> np.float32:
    if math.isnan(c):
        return np.float32(-np.inf)  # `v <= NaN` is False for every v: the cutoff is "below" everything
    with np.errstate(over="ignore"):
        t = np.float32(c)
    if float(t) > c:
        t = np.nextafter(t, np.float32(-np.inf), dtype=np.float32)
    return t


def native_thresholds(cutoffs, anv_dtype) -> np.ndarray:
    """float64 cutoffs -> native-type thresholds theta with (double(v) <= c) == (v <= theta), as uint64 slots."""
    n = len(cutoffs)
    raw = np.zeros(n, dtype=np.uint64)
    if anv_dtype == _lib.ANV_F32:
        th = np.array([_round_down_f32(float(c)) for c in cutoffs], dtype=np.float32)
        raw[:] = th.view(np.uint32).astype(np.uint64)
    elif anv_dtype == _lib.ANV_F64:
        th = np.array([(-np.inf if math.isnan(float(c)) else float(c)) for c in cutoffs], dtype=np.float64)
        raw[:] = th.view(np.uint64)
    else:
        lo, hi = (-(1 << 31), (1 << 31) - 1) if anv_dtype == _lib.ANV_I32 else (-(1 << 63), (1 << 63) - 1)
        vals = []
        for c in cutoffs:
            c = float(c)
            if math.isnan(c) or c == -math.inf:
                v = lo
            elif c == math.inf:
                v = hi
            else:
                v = min(max(math.floor(c), lo), hi)
            vals.append(v)
        if anv_dtype == _lib.ANV_I32:
            raw[:] = np.array(vals, dtype=np.int64).astype(np.int32).view(np.uint32).astype(np.uint64)
        else:
            raw[:] = np.array(vals, dtype=np.int64).view(np.uint64)
    return raw


class BinModel:
    """Host description of the binning of a set of columns + its device image."""

    def __init__(self, frame: ColumnFrame, names, cutoffs, lo_hi=None):
        self.names = list(names)
        self.cutoffs = [list(map(float, c)) for c in cutoffs]
        self.max_bins = max((len(c) + 1 for c in self.cutoffs), default=2)
        specs = np.zeros(len(self.names), dtype=_SPEC_DT)
        raws, off = [], 0
        for i, (nme, cut) in enumerate(zip(self.names, self.cutoffs)):
            col = frame.column(nme)
            raw = native_thresholds(cut, col.anv_dtype)
            mode, lo, inv_w = 0, 0.0, 0.0
            if lo_hi is not None and lo_hi[i] is not None and col.anv_dtype in (_lib.ANV_F32, _lib.ANV_F64):
                mn, mx = lo_hi[i]
                nb = len(cut) + 1
                w = (mx - mn) / nb
                th = raw.astype(np.uint32).view(np.float32).astype(np.float64) if col.anv_dtype == _lib.ANV_F32 \
                    else raw.view(np.float64)
                ulp = np.spacing(np.float32(max(abs(mn), abs(mx)))) if col.anv_dtype == _lib.ANV_F32 \
                    else np.spacing(max(abs(mn), abs(mx)))
                ok = (w > 0 and math.isfinite(w) and np.all(np.isfinite(th)) and np.all(np.diff(th) > 0)
                      and w >= 8 * float(ulp) and nb < (1 << 20))
                if ok:
                    mode, lo, inv_w = 1, mn, 1.0 / w
            specs[i] = (len(cut) + 1, mode, lo, inv_w, off)
            raws.append(raw)
            off += len(raw)
        self.specs_host = specs
        self.cuts_host = np.concatenate(raws) if raws else np.zeros(1, np.uint64)
        self._dev = None

    def device(self):
        if self._dev is None:
            self._dev = (_to_dev(self.specs_host), _to_dev(self.cuts_host if self.cuts_host.size else np.zeros(1, np.uint64)))
        return self._dev


# ---- K2 / fused / assign ---------------------------------------------------------------------

def histogram(frame: ColumnFrame, model: BinModel):
    """-> uint64 ndarray [n_cols, max_bins + 1]: slot 0 = nulls, slot b = rows in bin b."""
    global launch_count
    _lib.require_cuda()
    L = _lib.lib()
    n = len(model.names)
    stride = model.max_bins + 1
    if n == 0:
        return np.zeros((0, stride), np.uint64)
    desc, keep = frame.descriptors(model.names)
    specs, cuts = model.device()
    counts = _dev_bytes(n * stride * 8)
    _lib.check(L.anv_hist(desc.data_ptr(), specs.data_ptr(), cuts.data_ptr(), n, frame.n_rows, counts.data_ptr(),
                          stride, _stream()), "anv_hist")
    launch_count += 1
    return counts.cpu().numpy().view(np.uint64).reshape(n, stride).copy()


def moments_histogram(frame: ColumnFrame, model: BinModel):
    """Moments AND histogram of `model.names` in ONE read of the frame."""
    global launch_count
    _lib.require_cuda()
    L = _lib.lib()
    n = len(model.names)
    stride = model.max_bins + 1
    if n == 0:
        return np.zeros(0, dtype=_MOM_DT), np.zeros((0, stride), np.uint64)
    desc, keep = frame.descriptors(model.names)
    specs, cuts = model.device()
    counts = _dev_bytes(n * stride * 8)
    ws_bytes = L.anv_moments_workspace_bytes(n, frame.n_rows)
    ws = _dev_bytes(ws_bytes)
    out = _dev_bytes(n * _MOM_DT.itemsize)
    _lib.check(L.anv_moments_hist(desc.data_ptr(), specs.data_ptr(), cuts.data_ptr(), n, frame.n_rows, out.data_ptr(),
                                  counts.data_ptr(), stride, ws.data_ptr(), ws_bytes, _stream()), "anv_moments_hist")
    launch_count += 2
    return (out.cpu().numpy().view(_MOM_DT).copy(),
            counts.cpu().numpy().view(np.uint64).reshape(n, stride).copy())


def bin_assign(frame: ColumnFrame, model: BinModel):
    """-> int32 CUDA tensor [n_cols, stride] of bin ids (0 = null row); stride >= n_rows."""
    global launch_count
    torch = _lib.require_cuda()
    L = _lib.lib()
    n = len(model.names)
    stride = (frame.n_rows + 3) // 4 * 4
    out = torch.empty((max(n, 1), max(stride, 4)), dtype=torch.int32, device="cuda")
    if n == 0 or frame.n_rows == 0:
        return out[:n, :frame.n_rows]
    desc, keep = frame.descriptors(model.names)
    specs, cuts = model.device()
    _lib.check(L.anv_bin_assign(desc.data_ptr(), specs.data_ptr(), cuts.data_ptr(), n, frame.n_rows, model.max_bins,
                                out.data_ptr(), out.stride(0), _stream()), "anv_bin_assign")
    launch_count += 1
    return out[:, :frame.n_rows]


def code_counts(frame: ColumnFrame, names):
    """Dictionary-code histograms of string columns -> list of uint64 arrays [cardinality + 1]
    (slot 0 = nulls).  Columns are grouped by cardinality class so each launch sizes its
    shared-memory histogram for its own group."""
    global launch_count
    _lib.require_cuda()
    L = _lib.lib()
    out = {}
    groups = {}
    for nme in names:
        card = max(len(frame.column(nme).dictionary), 1)
        cls = 0 if card + 1 <= 40 else (1 if card + 1 <= 10240 else 2)
        groups.setdefault(cls, []).append(nme)
    for cls, grp in groups.items():
        cards = np.array([max(len(frame.column(g).dictionary), 1) for g in grp], dtype=np.int32)
        stride = int(cards.max()) + 1
        desc, keep = frame.descriptors(grp)
        dcards = _to_dev(cards)
        counts = _dev_bytes(len(grp) * stride * 8)
        _lib.check(L.anv_hist_codes(desc.data_ptr(), dcards.data_ptr(), len(grp), frame.n_rows, counts.data_ptr(),
                                    stride, _stream()), "anv_hist_codes")
        launch_count += 1
        h = counts.cpu().numpy().view(np.uint64).reshape(len(grp), stride)
        for i, g in enumerate(grp):
            out[g] = h[i, :cards[i] + 1].copy()
    return [out[n] for n in names]


# ---- K3 ---------------------------------------------------------------------------------

def drift_reduce(src_counts, tgt_counts, kinds, n_src, n_tgt, src_p=None):
    """Lists (per column) of aligned uint64 count arrays (slot 0 = nulls) -> structured array
    with psi/hd/jsd/ks/n_rows.  src_p: list of float64 arrays (NaN = key absent) instead of
    src_counts when the source comes from a saved model."""
    global launch_count
    _lib.require_cuda()
    L = _lib.lib()
    n = len(tgt_counts)
    if n == 0:
        return np.zeros(0, dtype=_DRIFT_DT)
    n_slots = np.array([len(t) for t in tgt_counts], dtype=np.int32)
    stride = int(n_slots.max())
    T = np.zeros((n, stride), np.uint64)
    for i, t in enumerate(tgt_counts):
        T[i, :len(t)] = t
    dT = _to_dev(T)
    if src_p is None:
        S = np.zeros((n, stride), np.uint64)
        for i, s in enumerate(src_counts):
            S[i, :len(s)] = s
        dS, dP, is_p = _to_dev(S), None, 0
    else:
        Pm = np.full((n, stride), np.nan, np.float64)
        for i, s in enumerate(src_p):
            Pm[i, :len(s)] = s
        dS, dP, is_p = None, _to_dev(Pm), 1
    dslots, dkind = _to_dev(n_slots), _to_dev(np.asarray(kinds, dtype=np.int32))
    out = _dev_bytes(n * _DRIFT_DT.itemsize)
    _lib.check(L.anv_drift_reduce(dS.data_ptr() if dS is not None else None, dT.data_ptr(),
                                  dP.data_ptr() if dP is not None else None, is_p, dslots.data_ptr(), dkind.data_ptr(),
                                  n, stride, int(n_src), int(n_tgt), out.data_ptr(), _stream()), "anv_drift_reduce")
    launch_count += 1
    return out.cpu().numpy().view(_DRIFT_DT).copy()
