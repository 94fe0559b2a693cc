import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from anovos_b200 import synth, profile, frame as F, engine
import bench
rows, cols = 10_000_000, 50
src = synth.device_frame(rows, cols)
host = bench.host_copy(src, torch)
del src
def run(group, pre=True, n=4):
    ts=[]
    for i in range(n):
        torch.cuda.synchronize(); t0=time.perf_counter()
        fr = F.ColumnFrame.from_tensors(host, n_rows=rows)
        if pre: profile.prefetch(fr, group=group)
        t1=time.perf_counter(); l0=engine.launch_count
        bench.stats_step(fr, keep_cache=pre)
        torch.cuda.synchronize(); t2=time.perf_counter()
        ts.append((round((t1-t0)*1e3,1), round((t2-t1)*1e3,1), engine.launch_count-l0))
        del fr
    return ts
print("pinned?", [ (v[0] if isinstance(v,tuple) else v).is_pinned() for v in list(host.values())[:2]])
x = list(host.values())[0]; x = x[0] if isinstance(x, tuple) else x
print("numpy roundtrip pinned?", torch.from_numpy(x.numpy()).is_pinned())
# raw copy bandwidth
torch.cuda.synchronize(); t0=time.perf_counter()
devs=[]
for v in host.values():
    t = v[0] if isinstance(v, tuple) else v
    devs.append(t.cuda(non_blocking=True))
torch.cuda.synchronize(); print("raw H2D of all columns: %.1f ms"%((time.perf_counter()-t0)*1e3)); del devs
print("direct     ", run(0, pre=False))
for g in (50, 25, 10, 5):
    print("prefetch g=%d"%g, run(g))
