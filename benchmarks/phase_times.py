#!/usr/bin/env python
"""Where the time of the sharded config-5 functions goes (torchrun, rank 0 prints)."""
import os
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from anemoi_transform_b200 import _cabi, spatial  # noqa: E402
from anemoi_transform_b200 import distributed as atd  # noqa: E402
from anemoi_transform_b200 import synthetic as syn  # noqa: E402
from anemoi_transform_b200.device import KnnIndex, compact_mask  # noqa: E402

rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
_cabi.load(check_device=True)
torch.cuda.set_device(local)
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lam = syn.rotated_lam(1000, 1000, 0.018, 60.0, 10.0)
glob = syn.octahedral(1280)


class T:
    def __init__(self):
        self.rows = []
        self.t = None

    def mark(self, name):
        torch.cuda.synchronize()
        now = time.perf_counter()
        if self.t is not None:
            self.rows.append((name, now - self.t))
        self.t = now


for rep in range(2):
    t = T()
    if ws > 1:
        dist.barrier()
    t.mark("start")
    gx = atd.latlon_to_xyz_device(*glob)
    t.mark("xyz global (sharded host trig + all-gather)")
    gi = KnnIndex(gx)
    t.mark("bucket build 6.6M")
    lx = atd.latlon_to_xyz_device(*lam)
    t.mark("xyz lam")
    lo, hi = atd.shard_range(gi.n, rank, ws)
    d = atd.all_reduce_min(gi.min_nn_distance(lo, hi - lo), device="cuda")
    t.mark("resolution (self 2-NN of my slice + all-reduce MIN)")
    lo, hi = atd.shard_range(int(lx[0].shape[0]), rank, ws)
    mark = gi.ball_mark(tuple(a[lo:hi] for a in lx), d)
    t.mark("ball mark")
    mark = atd.all_reduce_or(mark)
    t.mark("all-reduce OR")
    idx = compact_mask(mark).cpu().numpy()
    t.mark("compact + D2H")
    if rank == 0 and rep == 1:
        for name, s in t.rows:
            print(f"  {name:55s} {s * 1e3:8.2f} ms")
        t0 = time.perf_counter()
        x = spatial.latlon_to_xyz(*glob)
        t1 = time.perf_counter()
        KnnIndex(x)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"  single: host trig of all 6.6M {1e3 * (t1 - t0):.2f} ms, build from host arrays {1e3 * (t2 - t1):.2f} ms")
if ws > 1:
    dist.destroy_process_group()
