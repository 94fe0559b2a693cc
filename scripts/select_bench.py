"""Per-call device time of the radix select (resident and chunked) on the c2 frame."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from anovos_b200 import engine, profile, synth
from anovos_b200.partitioned import PartitionedFrame

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 50
fr = synth.device_frame(rows, cols)
names = fr.columns
mom = engine.moments(fr, names)
rk = np.array([engine.quantile_ranks(int(m["n_valid"]), profile.SUMMARY_PROBS) for m in mom], dtype=np.int64)
pf = PartitionedFrame.from_frame(fr, rows // 4)
for label, f in (("resident", fr), ("4 chunks", pf)):
    engine.select_ranks(f, names, rk)
    torch.cuda.synchronize()
    engine.timer = engine.KernelTimer()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        v = engine.select_ranks(f, names, rk)
    e1.record()
    torch.cuda.synchronize()
    print(label, "wall %.2f ms/call" % (e0.elapsed_time(e1) / 3), {k: round(x["ms"] / 3, 3) for k, x in engine.timer.totals().items()})
    engine.timer = None
