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


def icon_grid(path: str, refinement_level_c: int | None = None) -> tuple[np.ndarray, np.ndarray]:
    """Cell-centre latitudes / longitudes (degrees) of an ICON grid file, optionally only the
    cells with `refinement_level_c <= refinement_level_c` (reference `grids/icon.py:22-53`).

    ICON grid files are NetCDF (`clat`, `clon` in radians, `refinement_level_c`); they are read
    with xarray when it is installed.  A `.npz` with the same three arrays is accepted as well,
    so the filter can be used (and tested) where xarray / netCDF are not available."""
    if str(path).endswith(".npz"):
        ds = np.load(path)
        clat, clon, level = ds["clat"], ds["clon"], ds["refinement_level_c"] if "refinement_level_c" in ds.files else None
    else:
        try:
            import xarray as xr
        except ImportError as e:  # pragma: no cover - depends on the environment
            raise ImportError(f"reading the ICON grid {path} needs xarray (or provide clat / clon / refinement_level_c as .npz)") from e
        ds = xr.open_dataset(path)
        clat, clon, level = ds.clat.values, ds.clon.values, ds.refinement_level_c.values
    if refinement_level_c is not None:
        if level is None:
            raise ValueError(f"{path} has no refinement_level_c")
        keep = level <= refinement_level_c
        clat, clon = clat[keep], clon[keep]
    return np.rad2deg(clat), np.rad2deg(clon)
