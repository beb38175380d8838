"""Oracle: the tabular kNN filters on the §8(f) "next" list.

TEST INFRASTRUCTURE — see oracle/__init__.py.

Reference: `filters/tabular/support/superob.py:48-74` (`assign_nearest_grid`: cKDTree of the
grid's [lat, lon] pairs, nearest grid point and distance of every observation in the flat
plane, time slots by `searchsorted`) and `filters/tabular/superob.py:62-106` (`SuperOb.forward`:
group by (grid_index, …), mean of the value columns, nearest-observation columns by
`idxmin` of the distance).  scipy and pandas are what the reference itself calls; this module
makes the same calls.  PINNED against the imported reference in tests/test_oracle_tabular.py.
"""

from __future__ import annotations

import numpy as np
import pandas as pd
from scipy.spatial import cKDTree


def assign_nearest_grid(df: pd.DataFrame, grid_points: np.ndarray, time_slot_len: int) -> pd.DataFrame:
    df = df.copy()
    time_grid = pd.date_range(df["date"].min(), df["date"].max(), freq=f"{time_slot_len}s")
    temporal = np.clip(np.searchsorted(time_grid, df["date"], side="right") - 1, 0, None)
    distances, spatial = cKDTree(grid_points).query(df[["latitude", "longitude"]])
    return df.assign(grid_index=spatial + len(grid_points) * temporal, spatial_index=spatial, distance=distances)


def superob(df: pd.DataFrame, grid_points: np.ndarray, timeslot_length: int, take_nearest=(), groupby=()) -> pd.DataFrame:
    if len(df) == 0:
        return df
    df = df.dropna(subset=["date", "latitude", "longitude"])
    if len(df) == 0:
        return df
    df = assign_nearest_grid(df, grid_points, timeslot_length)
    keys = ["grid_index", *groupby]
    to_average = [c for c in df.columns if c not in (set(keys) | set(take_nearest))]
    averaged = df.groupby(keys, observed=True, sort=False)[to_average].mean()
    nearest_idx = df.groupby(keys, observed=True, sort=False)["distance"].idxmin()
    nearest = df.loc[nearest_idx, list(take_nearest) + keys].set_index(keys)
    averaged = averaged[~averaged.index.duplicated(keep="first")]
    nearest = nearest[~nearest.index.duplicated(keep="first")]
    out = pd.concat([averaged, nearest], axis=1, join="inner").reset_index()
    return out.drop(columns=["grid_index", "distance"], errors="ignore").sort_values("date")


def synthetic_observations(n: int, seed: int = 0) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    start = np.datetime64("2024-01-01T00:00:00")
    return pd.DataFrame(
        {
            "date": start + rng.integers(0, 6 * 3600, n).astype("timedelta64[s]"),
            "latitude": rng.uniform(-89.0, 89.0, n),
            "longitude": rng.uniform(-180.0, 180.0, n),
            "reporttype": rng.integers(0, 3, n),
            "obsvalue": rng.standard_normal(n) * 5 + 280,
            "height": rng.uniform(0, 2000, n),
        }
    )
