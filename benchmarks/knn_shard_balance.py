#!/usr/bin/env python
"""Config 2 on ONE GPU: how long does each rank's shard of the queries take alone, for
contiguous shards (rank r = queries [r·n/N, (r+1)·n/N)) and interleaved ones (q[r::N])?
The sharded step is as slow as its slowest rank (max over ranks), so this says what the
N-GPU query kernel can reach before any exchange."""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from anemoi_transform_b200 import _cabi, spatial  # noqa: E402
from anemoi_transform_b200 import synthetic as syn  # noqa: E402
from anemoi_transform_b200.device import KnnIndex, to_device_f64  # noqa: E402

_cabi.load(check_device=True)
s_xyz = spatial.latlon_to_xyz(*syn.regular_latlon(0.25))
t_xyz = spatial.latlon_to_xyz(*syn.n320_like())
knn = KnnIndex(s_xyz)
q_all = tuple(to_device_f64(a) for a in t_xyz)
nq = int(q_all[0].shape[0])


def timed(q, iters=50):
    for _ in range(5):
        knn.query(q, k=1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        knn.query(q, k=1)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


out = {"n_queries": nq, "all_ms": timed(q_all)}
for n in (2, 4, 8):
    per = -(-nq // n)
    cont = [timed(tuple(a[r * per : (r + 1) * per].contiguous() for a in q_all)) for r in range(n)]
    inter = [timed(tuple(a[r::n].contiguous() for a in q_all)) for r in range(n)]
    out[f"shards_{n}"] = {"contiguous_ms": cont, "interleaved_ms": inter, "contiguous_max": max(cont), "interleaved_max": max(inter)}
print(json.dumps(out, indent=1))
