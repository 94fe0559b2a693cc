"""World-size-2 gloo test of the N>1 host path: column sharding + the summary all_gather
(the only exchange of the hot path).  Runs on CPU."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import pandas as pd
    import torch.distributed as dist
    from anovos_b200 import parallel
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = ["c%02d" % i for i in range(7)]                 # 7 columns over 2 ranks: 4 + 3 (ragged)
    mine = parallel.shard_columns(names, rank, world)
    f1 = pd.DataFrame({"attribute": mine, "mean": [float(n[1:]) for n in mine], "mode": ["x"] * len(mine)})
    f2 = pd.DataFrame({"attribute": mine, "skewness": [10.0 + float(n[1:]) for n in mine]})
    m, fields = parallel.frames_to_matrix([f1, f2])
    parts = parallel.gather_summaries(m)
    h = parallel.gather_summaries_async(m, 4)                 # fixed-shape, non-blocking variant
    parts2 = [p[~np.isnan(p[:, 0])] for p in h.result()]
    assert all(np.array_equal(a, b) for a, b in zip(parts, parts2))
    dist.destroy_process_group()
    ret[rank] = (mine, fields, [p.tolist() for p in parts])


def test_column_shards_and_summary_gather_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    a, b = ret[0], ret[1]
    assert a[0] == ["c00", "c01", "c02", "c03"] and b[0] == ["c04", "c05", "c06"]
    assert a[1] == ["mean", "skewness"]
    assert a[2] == b[2]                                       # every rank ends with the same global table
    full = np.concatenate([np.array(p) for p in a[2]])
    assert np.array_equal(full[:, 0], np.arange(7.0)) and np.array_equal(full[:, 1], 10.0 + np.arange(7.0))
