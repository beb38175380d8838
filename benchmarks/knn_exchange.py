#!/usr/bin/env python
"""Config 2's sharded query under torchrun: the three ways of getting every rank the full index
array — stores from inside the search kernels, one coalesced broadcast kernel after the search,
NCCL all-gather in place — timed with CUDA events (max over ranks), results checked against a
single-GPU query."""
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from anemoi_transform_b200 import _cabi, spatial  # noqa: E402
from anemoi_transform_b200 import synthetic as syn  # noqa: E402
from anemoi_transform_b200.device import KnnIndex  # noqa: E402
from anemoi_transform_b200.distributed import ShardedKnnQuery  # noqa: E402

rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
_cabi.load(check_device=True)
torch.cuda.set_device(local)
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s_xyz = spatial.latlon_to_xyz(*syn.regular_latlon(0.25))
t_xyz = spatial.latlon_to_xyz(*syn.n320_like())
knn = KnnIndex(s_xyz)
out = {"world_size": ws}
for name, kw in (("peer_inline", dict(mode="peer", exchange="inline")), ("peer_bulk", dict(mode="peer", exchange="bulk")), ("nccl_in_place", dict(mode="nccl"))):
    if ws == 1 and name != "peer_bulk":
        continue
    q = ShardedKnnQuery(knn, t_xyz, k=1, **kw)
    for _ in range(5):
        q.step()
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        q.step()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 100], device="cuda", dtype=torch.float64)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = q.check_against_single_gpu() if ws > 1 else True
    out[name] = {"ms": float(t.item()), "identical_to_single_gpu": ok}
    q.close()
if rank == 0:
    print(json.dumps(out), flush=True)
if ws > 1:
    dist.destroy_process_group()
