"""cProfile of one streamed drift step (bench.py c4 shape at reduced rows): where the host time goes."""
import cProfile
import pstats
import sys
import tempfile

sys.path.insert(0, ".")
import torch
import anovos.data_analyzer.stats_generator as sg
import anovos.drift_stability.drift_detector as dd
from anovos_b200 import synth

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 200
chunk = int(float(sys.argv[3])) if len(sys.argv) > 3 else 2_500_000
tmp = tempfile.mkdtemp()


def step():
    src = synth.partitioned_frame(rows, cols, chunk, seed=42, cat_every=4)
    tgt = synth.partitioned_frame(rows, cols, chunk, seed=43, shifted=True, cat_every=4)
    r = [dd.statistics(None, tgt, src, method_type="all", use_sampling=False, source_path=tmp).toPandas()]
    for f in (src, tgt):
        r += [sg.measures_of_counts(None, f).toPandas(), sg.measures_of_shape(None, f).toPandas()]
    return r


step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
