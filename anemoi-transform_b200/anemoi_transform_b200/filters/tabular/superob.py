"""`superob` — reference `filters/tabular/superob.py:20-106` with
`support/superob.py:48-74` (`assign_nearest_grid`).

Observations are binned in space (nearest point of the output grid, searched in the flat
latitude / longitude plane) and time (slots of `timeslot_length` seconds from the first
observation); per bin the numeric columns are averaged and the `columns_to_take_nearest`
columns are taken from the observation closest to the grid point.

The spatial search — `cKDTree(grid).query(obs)` on the host in the reference — is the device
kNN (`assign_to_grid.nearest_in_plane`: distances bitwise cKDTree's, indices equal off exact
ties); binning and aggregation are pandas on the host, as in the reference (they are not on
the hot path and depend on pandas' own group-by semantics).
"""

from __future__ import annotations

from typing import Any

import numpy as np

from ...filter import Filter
from . import filter_registry
from .assign_to_grid import define_grid, define_healpix_grid, nearest_in_plane


def assign_nearest_grid(df: Any, grid_points: np.ndarray, time_slot_len: int) -> Any:
    """→ a copy of `df` with `spatial_index` (nearest grid point), `distance` (to it, in
    degrees of the lat-lon plane) and `grid_index` (= spatial_index + n_grid · time slot)."""
    import pandas as pd

    slots = pd.date_range(df["date"].min(), df["date"].max(), freq=f"{time_slot_len}s")
    # the slot an observation falls into: the last slot start that is <= its date
    slot_of = np.maximum(np.searchsorted(slots, df["date"], side="right") - 1, 0)
    distance, nearest = nearest_in_plane(grid_points, df[["latitude", "longitude"]].to_numpy())
    return df.copy().assign(grid_index=nearest + len(grid_points) * slot_of, spatial_index=nearest, distance=distance)


@filter_registry.register("superob")
class SuperOb(Filter):
    """Aggregate observations into the cells of `grid` and time slots of `timeslot_length` s."""

    def __init__(self, *, grid: str, timeslot_length: int, columns_to_take_nearest: list[str] | None = None, columns_to_groupby: list[str] | None = None):
        self.grid = grid
        self.timeslot_length = timeslot_length
        self.columns_to_take_nearest = list(columns_to_take_nearest or [])
        self.columns_to_groupby = list(columns_to_groupby or [])

    def forward(self, df: Any) -> Any:
        import pandas as pd

        if self.grid == "native" or len(df) == 0:
            return df
        grid_points = define_healpix_grid(int(self.grid[1:])) if self.grid[0] == "h" else define_grid(self.grid)
        df.dropna(subset=["date", "latitude", "longitude"], inplace=True)  # in place, like the reference
        if len(df) == 0:
            return df
        binned = assign_nearest_grid(df, grid_points, self.timeslot_length)

        keys = ["grid_index", *self.columns_to_groupby]
        not_averaged = set(keys) | set(self.columns_to_take_nearest)
        bins = binned.groupby(keys, observed=True, sort=False)
        means = bins[[c for c in binned.columns if c not in not_averaged]].mean()
        closest = binned.loc[bins["distance"].idxmin(), self.columns_to_take_nearest + keys].set_index(keys)
        means = means[~means.index.duplicated(keep="first")]
        closest = closest[~closest.index.duplicated(keep="first")]
        out = pd.concat([means, closest], axis=1, join="inner").reset_index()
        return out.drop(columns=["grid_index", "distance"], errors="ignore").sort_values("date")
