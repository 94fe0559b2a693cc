"""Kernel micro-benchmark: K1 / K2 / fused on synthetic f32 columns resident in HBM.
Usage: python scripts/kbench.py [rows] [cols] [null_mode]"""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from anovos_b200 import _lib, engine
from anovos_b200.frame import ColumnFrame

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 50
null_mode = sys.argv[3] if len(sys.argv) > 3 else "mixed"   # none | all | mixed
L = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
data = {}
t0 = time.time()
for c in range(cols):
    x = torch.empty(rows, dtype=torch.float32, device="cuda")
    fam = c % 4
    rate = {"none": 0.0, "all": 0.02, "mixed": [0.0, 0.001, 0.02, 0.3][c % 4]}[null_mode]
    v = torch.zeros((rows + 31) // 32, dtype=torch.int32, device="cuda") if rate > 0 else None
    a, b = [(5.0 + c, 1.0 + 0.1 * c), (0.0, 0.75), (-3.0 - c, 7.0 + c), (0.0, 2.0)][fam]
    _lib.check(L.anv_synth_f32(x.data_ptr(), v.data_ptr() if v is not None else None, rows, 42, c, fam, a, b, rate, st))
    data["c%03d" % c] = (x, v) if v is not None else x
torch.cuda.synchronize()
print("generated %.2f GB in %.2fs" % (rows * cols * 4 / 1e9, time.time() - t0), flush=True)
fr = ColumnFrame.from_tensors(data)
names = fr.columns
mom = engine.moments(fr, names)
cuts = [[float(mom["min"][i]) + j * ((float(mom["max"][i]) - float(mom["min"][i])) / 10) for j in range(1, 10)] for i in range(cols)]
model = engine.BinModel(fr, names, cuts, [(float(mom["min"][i]), float(mom["max"][i])) for i in range(cols)])
print("modes", np.bincount(model.specs_host["mode"]))
nbytes = rows * cols * 4 + sum((rows + 7) // 8 for c in range(cols) if isinstance(data["c%03d" % c], tuple))


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


desc, keep = fr.descriptors(names)
specs, dcuts = model.device()
ws_bytes = L.anv_moments_workspace_bytes(cols, rows)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
out = torch.empty(cols * 64, dtype=torch.uint8, device="cuda")
counts = torch.empty(cols * 11 * 8, dtype=torch.uint8, device="cuda")
res = {}
res["moments"] = timeit(lambda: _lib.check(L.anv_moments(desc.data_ptr(), cols, rows, out.data_ptr(), ws.data_ptr(), ws_bytes, st)))
res["hist"] = timeit(lambda: _lib.check(L.anv_hist(desc.data_ptr(), specs.data_ptr(), dcuts.data_ptr(), cols, rows, counts.data_ptr(), 11, st)))
res["fused"] = timeit(lambda: _lib.check(L.anv_moments_hist(desc.data_ptr(), specs.data_ptr(), dcuts.data_ptr(), cols, rows, out.data_ptr(), counts.data_ptr(), 11, ws.data_ptr(), ws_bytes, st)))
a = torch.empty(rows * cols // 2, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
res["copy(torch)"] = timeit(lambda: b.copy_(a))
for k, (med, mn) in res.items():
    by = nbytes if k != "copy(torch)" else a.numel() * 8
    print("%-12s median %.3f ms  min %.3f ms  -> %.0f GB/s (median)  %.3g rows*cols/s" % (k, med, mn, by / med / 1e6, rows * cols / med * 1e3))
print(json.dumps({"rows": rows, "cols": cols, "nulls": null_mode, "bytes": nbytes, **{k: v[0] for k, v in res.items()}}))
