#!/usr/bin/env python
"""End-to-end throughput of the drop-in regrid filter on ordinary numpy fields (config 3 grids):

    FieldList of pageable float32 fields -> RegridFilter.forward -> to_numpy() of every output

plus the legs it is made of (upload only, download only), for a sweep of staging-thread
counts.  Writes one JSON document to stdout.

    python benchmarks/plugin_e2e.py [--fields 3120] [--threads 4,8,12,15] [--reps 3]
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fields", type=int, default=3120)
    ap.add_argument("--threads", default="0")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--legs", action="store_true", help="also time upload-only and download-only")
    ap.add_argument("--profile", action="store_true", help="cProfile one plugin call (stderr)")
    args = ap.parse_args()

    import torch

    from anemoi_transform_b200 import ekd
    from anemoi_transform_b200 import synthetic as syn
    from anemoi_transform_b200.device import DeviceBatch, HostIO
    from anemoi_transform_b200.filters import create_filter_by_name

    t_lat, t_lon = syn.n320_like()
    s_lat, s_lon = syn.regular_latlon(0.25)
    d, i, p, shape = syn.bilinear_matrix(0.25, t_lat, t_lon)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "m.npz")
    syn.save_regrid_npz(path, d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    n_tgt, n_src = shape
    rng = np.random.default_rng(0)
    base = rng.standard_normal((64, n_src), dtype=np.float32)
    # every field its own pageable array (np.array copies), as a decoder would deliver them
    values = [np.array(base[k % 64]) for k in range(args.fields)]
    fl = ekd.from_source("list-of-dicts", [dict(param="t", levelist=850, step=k, values=v, latitudes=s_lat, longitudes=s_lon) for k, v in enumerate(values)])
    regrid = create_filter_by_name("regrid", matrix=path)
    in_gb, out_gb = args.fields * n_src * 4 / 1e9, args.fields * n_tgt * 4 / 1e9
    results = {"fields": args.fields, "host_cpus": os.cpu_count(), "in_gb": in_gb, "out_gb": out_gb, "runs": []}

    def plugin_once():
        t0 = time.perf_counter()
        out = regrid.forward(fl)
        t1 = time.perf_counter()
        arrays = [f.to_numpy() for f in out]
        t2 = time.perf_counter()
        return t2 - t0, t1 - t0, arrays

    for nt in [int(x) for x in args.threads.split(",")]:
        HostIO._engines.clear()
        if nt > 0:
            os.environ["AT_B200_COPY_THREADS"] = str(nt)
        else:
            os.environ.pop("AT_B200_COPY_THREADS", None)
        io = HostIO.get()
        run = {"threads": io.n_threads, "nontemporal": io.nontemporal}
        cold = plugin_once()[0]  # warm-up: pins the staging slots and grows the page-locked pool
        run["first_call_s"] = cold
        best = None
        for _ in range(args.reps):
            total, fwd, arrays = plugin_once()
            if best is None or total < best[0]:
                best = (total, fwd)
            del arrays
        run["plugin_s"], run["forward_s"] = best
        run["plugin_fields_per_s"] = args.fields / best[0]
        run["plugin_host_gb_per_s"] = (in_gb + out_gb) / best[0]
        if args.legs:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            batch = DeviceBatch.from_host_fields(values)
            torch.cuda.synchronize()
            run["upload_s"] = time.perf_counter() - t0
            run["upload_gb_per_s"] = in_gb / run["upload_s"]
            y = DeviceBatch(batch.data[:n_tgt], args.fields)
            t0 = time.perf_counter()
            y.prefetch()
            y._wait_all()
            run["download_s"] = time.perf_counter() - t0
            run["download_gb_per_s"] = out_gb / run["download_s"]
            del batch, y
        results["runs"].append(run)
        print(json.dumps(run), file=sys.stderr, flush=True)
    if args.profile:
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
        plugin_once()
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(18)
    # parity of the last run's path on a few fields
    from scipy.sparse import csr_array

    m = csr_array((d, i, p), shape=shape)
    out = regrid.forward(fl)
    results["parity"] = all(bool(np.array_equal((m @ values[k]).view(np.uint32), out[k].to_numpy().view(np.uint32))) for k in (0, args.fields // 2, args.fields - 1))
    print(json.dumps(results))


if __name__ == "__main__":
    main()
