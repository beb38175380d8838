#!/usr/bin/env python
"""Config 2: query time against the level-0 cell edge (multiples of the mean point spacing);
0 = the library's own choice.  Results must be identical for every cell size."""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from anemoi_transform_b200 import _cabi, spatial  # noqa: E402
from anemoi_transform_b200 import synthetic as syn  # noqa: E402
from anemoi_transform_b200.device import KnnIndex, to_device_f64  # noqa: E402

_cabi.load(check_device=True)
def ms(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


tq = tuple(to_device_f64(a) for a in spatial.latlon_to_xyz(*syn.n320_like()))
report = {}
for name, grid in (("0.25deg regular (crowded poles)", syn.regular_latlon(0.25)), ("O640 octahedral (even spacing)", syn.octahedral(640))):
    sx = spatial.latlon_to_xyz(*grid)
    spacing = float(np.sqrt(4 * np.pi / sx[0].size))
    out, ref = [], None
    for mult in (0.0, 0.8, 1.0, 1.25, 1.5, 2.0, 2.5, 3.0, 4.0):
        knn = KnnIndex(sx, cell_size=mult * spacing)
        row = {"cell_over_spacing": mult}
        for k in (1, 5, 12):
            row[f"k{k}_ms"] = round(ms(lambda: knn.query(tq, k=k)), 4)
        idx = knn.query(tq, k=1)[0]
        if ref is None:
            ref = idx
        row["identical"] = bool(torch.equal(idx, ref))
        out.append(row)
        knn.close()
    report[name] = out
print(json.dumps(report, indent=1))
