"""cProfile of the host side of one full stats_generator step (c2)."""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, ".")
from anovos_b200 import synth
import bench
fr = synth.device_frame(10_000_000, 50)
for _ in range(3):
    bench.stats_step(fr)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    bench.stats_step(fr)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    bench.stats_step(fr)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
