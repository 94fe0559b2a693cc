"""Profile only the sort-based mode/distinct path (ncu --profile-from-start off)."""
import sys
import torch
sys.path.insert(0, ".")
from anovos_b200 import engine, synth
rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 8
fr = synth.device_frame(rows, cols)
engine.sort_mode_distinct(fr, fr.columns)
torch.cuda.synchronize()
torch.cuda.profiler.start()
engine.sort_mode_distinct(fr, fr.columns)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
