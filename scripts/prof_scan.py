"""One launch of each scan-kernel variant (K1, K2, fused) + HLL + select inside a cudaProfiler range."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from anovos_b200 import engine, profile, synth
rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 50
fr = synth.device_frame(rows, cols)
names = fr.columns
mom = engine.moments(fr, names)
cuts = [[float(mom["min"][i]) + j * ((float(mom["max"][i]) - float(mom["min"][i])) / 10) for j in range(1, 10)] for i in range(cols)]
model = engine.BinModel(fr, names, cuts, [(float(mom["min"][i]), float(mom["max"][i])) for i in range(cols)])
ranks = np.array([engine.quantile_ranks(int(mom["n_valid"][i]), profile.SUMMARY_PROBS) for i in range(cols)])
def run():
    engine.moments(fr, names); engine.histogram(fr, model); engine.moments_histogram(fr, model)
    engine.hll_estimates(fr, names, 9); engine.select_ranks(fr, names, ranks)
run(); torch.cuda.synchronize()
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
