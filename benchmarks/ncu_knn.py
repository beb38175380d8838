#!/usr/bin/env python
"""Config 2's k = 1 query, two launches — the program `ncu --set full -k regex:knn_query -s 1 -c 1` captures."""
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
knn = KnnIndex(spatial.latlon_to_xyz(*syn.regular_latlon(0.25)))
q = tuple(to_device_f64(a) for a in spatial.latlon_to_xyz(*syn.n320_like()))
for _ in range(2):
    knn.query(q, k=1)
torch.cuda.synchronize()
