"""`assign_to_grid` — reference `filters/tabular/assign_to_grid.py:18-64`.

Adds `grid_index_{grid}` (index of the nearest grid point) and `distance` to a DataFrame of
observations.  The reference searches in the flat (latitude, longitude) plane with
`cKDTree(grid_points).query(obs[["latitude", "longitude"]])`; here the same 2-D Euclidean
search runs on the device kNN (`at_knn_create` / `at_knn_query`, the plane embedded at z = 0,
so `((dlat² + dlon²) + 0²)` is bitwise the 2-D distance cKDTree computes).
"""

from __future__ import annotations

from typing import Any

import numpy as np

from ...device import KnnIndex, require_cuda
from ...filter import Filter
from . import filter_registry


def define_grid(grid: str) -> np.ndarray:
    """Grid points as an (N, 2) array of [lat, lon] pairs, longitudes in (-180, 180]
    (reference `filters/tabular/support/superob.py:18-24`)."""
    from ...grids import lookup

    data = lookup(grid)
    lat = data["latitudes"]
    lon = data["longitudes"]
    lon = np.where(lon > 180, lon - 360, lon)
    return np.column_stack([lat, lon])


def define_healpix_grid(nside: int) -> np.ndarray:
    """HEALPix pixel centres as [lat, lon] pairs (reference superob.py:27-45); needs healpy."""
    import healpy as hp

    npix = hp.nside2npix(nside)
    theta, phi = hp.pix2ang(nside, np.arange(npix))
    lat = 90 - np.degrees(theta)
    lon = np.degrees(phi)
    lon = np.where(lon > 180, lon - 360, lon)
    return np.column_stack([lat, lon])


def nearest_in_plane(grid_points: np.ndarray, queries: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """`cKDTree(grid_points).query(queries)` for (N, 2) arrays → (distances, indices)."""
    require_cuda()
    grid_points = np.ascontiguousarray(grid_points, dtype=np.float64)
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    zeros_g, zeros_q = np.zeros(grid_points.shape[0]), np.zeros(queries.shape[0])
    knn = KnnIndex((np.ascontiguousarray(grid_points[:, 0]), np.ascontiguousarray(grid_points[:, 1]), zeros_g))
    try:
        idx, dist, _ = knn.query((np.ascontiguousarray(queries[:, 0]), np.ascontiguousarray(queries[:, 1]), zeros_q), k=1)
        return dist[:, 0].cpu().numpy(), idx[:, 0].cpu().numpy()
    finally:
        knn.close()


@filter_registry.register("assign_to_grid")
class AssignToGrid(Filter):
    """Adds a new column (``grid_index_{grid}``) to the DataFrame which represents the index of
    the nearest grid point, based on the latitude/longitude coordinates."""

    def __init__(self, *, grid: str):
        if not grid:
            raise ValueError("No grid specified.")
        self.grid = grid

    def forward(self, obs_df: Any) -> Any:
        if self.grid[0] == "h":
            grid_points = define_healpix_grid(int(self.grid[1:]))
        else:
            grid_points = define_grid(self.grid)
        distances, spatial_indices = nearest_in_plane(grid_points, obs_df[["latitude", "longitude"]].to_numpy())
        return obs_df.assign(**{f"grid_index_{self.grid}": spatial_indices}, distance=distances)
