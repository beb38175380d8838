"""Named grids.  The reference downloads `grid-{name}.npz` from get.ecmwf.int
(`grids/named.py:27-88`); offline, the grids this package can generate itself are served
from `synthetic` (octahedral O-grids are exact; see tests/test_grids)."""

from __future__ import annotations

import re
from typing import Any

import numpy as np

from . import synthetic


def lookup(name: Any) -> dict[str, np.ndarray]:
    if isinstance(name, (list, tuple)):
        if len(name) == 2 and float(name[0]) == float(name[1]):
            lat, lon = synthetic.regular_latlon(float(name[0]))
            return dict(latitudes=lat, longitudes=lon)
        raise ValueError(f"Invalid grid: {name}")
    key = str(name).lower()
    m = re.fullmatch(r"o(\d+)", key)
    if m:
        lat, lon = synthetic.octahedral(int(m.group(1)))
        return dict(latitudes=lat, longitudes=lon)
    raise ValueError(f"Unknown grid {name!r}: only octahedral O-grids and regular [d, d] grids are generated offline")
