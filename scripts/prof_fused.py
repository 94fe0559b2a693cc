"""One launch of the fused moments + histogram kernel per staging variant inside a cudaProfiler range (for ncu)."""
import os
import sys
import torch
sys.path.insert(0, ".")
from anovos_b200 import engine, synth
rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 50
fr = synth.device_frame(rows, cols)
names = fr.columns
mom = engine.moments(fr, names)
cuts = [[float(mom["min"][i]) + j * ((float(mom["max"][i]) - float(mom["min"][i])) / 10) for j in range(1, 10)] for i in range(cols)]
model = engine.BinModel(fr, names, cuts, [(float(mom["min"][i]), float(mom["max"][i])) for i in range(cols)])
def run():
    for flag in ("0", "1"):
        os.environ["ANV_FUSED_STAGED"] = flag
        engine.moments_histogram(fr, model)
run(); torch.cuda.synchronize()
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
