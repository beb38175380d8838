#!/usr/bin/env python
"""Repeatability and phases of thinning_mask (config 5) on one GPU."""
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from anemoi_transform_b200 import _cabi, spatial  # noqa: E402
from anemoi_transform_b200 import synthetic as syn  # noqa: E402
from anemoi_transform_b200.device import KnnIndex  # noqa: E402

_cabi.load(check_device=True)
lam = syn.rotated_lam(1000, 1000, 0.018, 60.0, 10.0)
glob = syn.octahedral(1280)
for rep in range(4):
    t0 = time.perf_counter()
    spatial.thinning_mask(*lam, *glob)
    torch.cuda.synchronize()
    print(f"thinning_mask call {rep}: {time.perf_counter() - t0:.3f} s", flush=True)
mask = spatial.cropping_mask(glob[0], glob[1], *spatial._crop_box(lam[0], lam[1], 2.0))
t0 = time.perf_counter(); gx = spatial.latlon_to_xyz(glob[0][mask], glob[1][mask]); lx = spatial.latlon_to_xyz(*lam); t1 = time.perf_counter()
print(f"host trig {t1 - t0:.3f} s for {gx[0].size} + {lx[0].size} points")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    index = KnnIndex(lx); torch.cuda.synchronize(); t1 = time.perf_counter()
    q = tuple(torch.from_numpy(a).cuda() for a in gx); torch.cuda.synchronize(); t2 = time.perf_counter()
    for part in (slice(None), slice(0, gx[0].size // 8), slice(7 * (gx[0].size // 8), None)):
        qq = tuple(a[part].contiguous() for a in q)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        idx, _, _ = index.query(qq, k=1); torch.cuda.synchronize(); t4 = time.perf_counter()
        print(f"  rep {rep}: build {t1 - t0:.3f} upload {t2 - t1:.3f} query[{part}] {t4 - t3:.4f} s", flush=True)
