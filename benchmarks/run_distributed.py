#!/usr/bin/env python
"""BASELINE config 5 under torchrun: the LAM masks with query points sharded over the ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        benchmarks/run_distributed.py [--small] [--out gpurun_out/config5_dist.json]

Every rank computes the sharded `global_on_lam_mask`, `thinning_mask`, `cutout_mask` and
`nearest_grid_points` (NCCL all-gather / all-reduce of the per-rank pieces); rank 0 also computes
them on one GPU and checks that the sharded results are identical.  Times are the max over ranks.
"""

from __future__ import annotations

import argparse
import json
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--out", default="gpurun_out/config5_dist.json")
    a = ap.parse_args()
    rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    _cabi.load(check_device=True)
    torch.cuda.set_device(local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lam = syn.rotated_lam(200, 240, 0.05, 55.0, 15.0) if a.small else syn.rotated_lam(1000, 1000, 0.018, 60.0, 10.0)
    glob = syn.octahedral(160 if a.small else 1280)

    def timed(fn):
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    res = {"world_size": ws, "lam_points": int(lam[0].size), "global_points": int(glob[0].size)}
    cases = {
        "global_on_lam_mask": (lambda: atd.global_on_lam_mask(*lam, *glob), lambda: spatial.global_on_lam_mask(*lam, *glob)),
        "thinning_mask": (lambda: atd.thinning_mask(*lam, *glob), lambda: spatial.thinning_mask(*lam, *glob)),
        "cutout_mask": (lambda: atd.cutout_mask(*lam, *glob), lambda: spatial.cutout_mask(*lam, *glob)),
        "nearest_grid_points": (lambda: atd.nearest_grid_points(*glob, *lam), lambda: spatial.nearest_grid_points(*glob, *lam)),
    }
    for name, (sharded, single) in cases.items():
        sharded()  # warm-up (bucket build allocations, NCCL channels)
        t, got = timed(sharded)
        res[f"{name}_sharded_s"] = t
        if rank == 0:
            t0 = time.perf_counter()
            want = single()
            res[f"{name}_single_gpu_s"] = time.perf_counter() - t0
            res[f"{name}_identical"] = bool(np.array_equal(got, want))
        if ws > 1:
            dist.barrier()
    if rank == 0:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(json.dumps(res, indent=1))
        print(json.dumps(res), flush=True)
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
