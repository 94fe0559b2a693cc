"""Wall time of each stats function in a full step vs the device time of the C calls it makes (c2)."""
import sys, time
import torch
sys.path.insert(0, ".")
from anovos_b200 import synth, engine
import anovos.data_analyzer.stats_generator as sg
fr = synth.device_frame(10_000_000, 50)
fns = ["measures_of_counts", "measures_of_centralTendency", "measures_of_cardinality", "measures_of_dispersion",
       "measures_of_percentiles", "measures_of_shape"]
def step(timed=False):
    fr._cache = {k: v for k, v in fr._cache.items() if isinstance(k, tuple) and k and k[0] == "desc"}
    out = {}
    for f in fns:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = getattr(sg, f)(None, fr).toPandas()
        torch.cuda.synchronize(); out[f] = (time.perf_counter() - t0) * 1e3
    return out
for _ in range(3): step()
acc = {f: 0.0 for f in fns}
engine.timer = engine.KernelTimer()
N = 10
for _ in range(N):
    for f, t in step().items(): acc[f] += t
kt = engine.timer.totals(); engine.timer = None
for f in fns: print("%-30s %.2f ms" % (f, acc[f] / N))
print("sum %.2f ms" % (sum(acc.values()) / N))
print({k: round(v["ms"] / N, 2) for k, v in kt.items()})
