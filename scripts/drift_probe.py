import cProfile, pstats, sys, time, tempfile
import torch
sys.path.insert(0, ".")
from anovos_b200 import synth, engine
import anovos.drift_stability.drift_detector as dd
rows, cols = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 50
src = synth.device_frame(rows, cols, seed=42)
tgt = synth.device_frame(rows, cols, seed=43, shifted=True)
d = tempfile.mkdtemp()
def run():
    src._cache = {k: v for k, v in src._cache.items() if isinstance(k, tuple) and k[0] == "desc"}
    tgt._cache = {k: v for k, v in tgt._cache.items() if isinstance(k, tuple) and k[0] == "desc"}
    return dd.statistics(None, tgt, src, method_type="all", use_sampling=False, source_path=d)
for _ in range(3): r = run()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): r = run()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("drift statistics: %.2f ms  -> %.3g rows*cols/s (one frame)" % (dt * 1e3, rows * cols / dt))
print(r.toPandas().head(8).to_string())
engine.timer = engine.KernelTimer(); run(); print(engine.timer.totals()); engine.timer = None
pr = cProfile.Profile(); pr.enable()
for _ in range(3): run()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
