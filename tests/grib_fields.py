"""A stand-in for earthkit-data's `GribField` (test infrastructure): a field that owns an encoded
GRIB message, exposes it through `message()`, and decodes it on the host in `to_numpy()` the
way ecCodes would (here: `oracle.grib.decode`).  Its metadata comes from a dict, like the other
test fields."""

from __future__ import annotations

import numpy as np

from anemoi_transform_b200 import ekd
from oracle import grib as ogrib


class GribMessageField(ekd.ArrayField):
    def __init__(self, message: bytes, n_points: int, metadata: dict, *, latitudes=None, longitudes=None):
        # ArrayField wants an array for its shape; the values are never read from it
        super().__init__(np.broadcast_to(np.float64(np.nan), (n_points,)), metadata, latitudes=latitudes, longitudes=longitudes)
        self._message = message
        self._n_points = n_points
        self.decodes = 0  # how often the host decoder ran

    @property
    def shape(self):
        return (self._n_points,)

    def message(self) -> bytes:
        return self._message

    def to_numpy(self, flatten: bool = False, dtype=None, index=None) -> np.ndarray:
        self.decodes += 1
        v = ogrib.decode(self._message, n_points=self._n_points)
        if dtype is not None:
            v = v.astype(dtype)
        if index is not None:
            v = v[index]
        return v

    @property
    def values(self):
        return self.to_numpy(flatten=True)
