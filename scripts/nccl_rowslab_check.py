"""Two (or more) GPUs, NCCL: the row-slab merge and the row-slab -> column-block exchange against the
single-frame result computed locally on every rank.  Launch with

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        scripts/nccl_rowslab_check.py

Prints "ROWSLAB-OK <world>" on rank 0; any mismatch raises."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

import anovos.data_analyzer.stats_generator as sg
import anovos.drift_stability.drift_detector as dd
from anovos_b200 import engine, parallel, synth
from anovos_b200.partitioned import PartitionedFrame, repartition_to_columns

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ROWS, COLS = 1_000_000, 12
per = ROWS // world // 32 * 32
r0 = rank * per
r1 = ROWS if rank == world - 1 else r0 + per


def mk(seed, a=0, b=ROWS):
    return synth.device_frame(b - a, COLS, seed=seed, cat_every=4, row0=a, shifted=seed != 42)


whole, twhole = mk(42), mk(43)
slab, tslab = mk(42, r0, r1), mk(43, r0, r1)
parts = PartitionedFrame.from_frame(slab, 131_072, group=True)
tparts = PartitionedFrame.from_frame(tslab, 131_072, group=True)
assert parts.count() == ROWS
mw, mp = engine.moments(whole, whole.columns), engine.moments(parts, whole.columns)
for f in ("n_valid", "n_nonzero", "min", "max"):
    assert np.array_equal(mw[f], mp[f], equal_nan=True), f
for f in ("mean", "m2", "m4"):
    assert np.allclose(mw[f], mp[f], rtol=1e-9, atol=0), f
for fn in ("measures_of_counts", "measures_of_percentiles", "measures_of_cardinality"):
    a, b = getattr(sg, fn)(None, whole).toPandas(), getattr(sg, fn)(None, parts).toPandas()
    assert a.equals(b), fn
kw = dict(method_type="all", use_sampling=False)
a = dd.statistics(None, twhole, whole, source_path="/tmp/anv_nccl_w%d" % rank, **kw).toPandas()
b = dd.statistics(None, tparts, parts, source_path="/tmp/anv_nccl_p%d" % rank, **kw).toPandas()
for m in ("PSI", "HD", "JSD", "KS"):
    assert np.allclose(a[m], b[m], rtol=1e-9, atol=0), m
assert list(a["flagged"]) == list(b["flagged"])
# exact mode on the row slabs directly (exchange + sort + all_gather under the hood)
cw_all = sg.measures_of_centralTendency(None, whole).toPandas()
cp_all = sg.measures_of_centralTendency(None, parts).toPandas()
assert cw_all.equals(cp_all), (cw_all, cp_all)
# the exchange over NVLink: slabs -> whole columns of this rank's block, then exact mode / distinct
mine = repartition_to_columns(slab, True)
names = parallel.shard_columns(slab.columns, rank, world)
assert mine.columns == names and mine.count() == ROWS
for n in names:
    d, v = mine.column(n).device()
    dw, vw = whole.column(n).device()
    assert torch.equal(d, dw), n
    assert (v is None and vw is None) or torch.equal(v.view(torch.int32), vw.view(torch.int32)), n
cw = sg.measures_of_centralTendency(None, whole, names).toPandas().sort_values("attribute").reset_index(drop=True)
cm = sg.measures_of_centralTendency(None, mine).toPandas().sort_values("attribute").reset_index(drop=True)
assert cw.equals(cm), (cw, cm)
dist.barrier()
if rank == 0:
    print("ROWSLAB-OK", world, flush=True)
dist.destroy_process_group()
