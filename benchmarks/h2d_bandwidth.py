#!/usr/bin/env python
"""Plain pinned-memory cudaMemcpyAsync bandwidth of the box, all ranks at once: the host-side
ceiling every end-to-end number of bench.py sits under.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        benchmarks/h2d_bandwidth.py [--out profiles/r02_h2d_bandwidth_Ngpu.json]

Each rank copies a 1 GiB pinned buffer to its GPU (H2D), back (D2H), and both at once on two
streams; times are CUDA events after a barrier, aggregate = sum over ranks of bytes / max time.
"""

from __future__ import annotations

import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=8)
    a = ap.parse_args()
    rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.mib << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d: bool, d2h: bool) -> float:
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(a.iters):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    out = {"world_size": ws, "host_cpus": os.cpu_count(), "buffer_mib": a.mib, "iters": a.iters}
    for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        run(h2d, d2h)
        t = torch.tensor([run(h2d, d2h)], device="cuda", dtype=torch.float64)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_dir = a.iters * n / float(t.item()) / 1e9
        out[name] = {"per_rank_GBps_per_direction": per_dir, "aggregate_GBps": per_dir * ws * (2 if (h2d and d2h) else 1)}
    if rank == 0:
        print(json.dumps(out), flush=True)
        if a.out:
            with open(a.out, "w") as f:
                json.dump(out, f, indent=1)
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
