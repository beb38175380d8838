"""Oracle: interpolation-matrix construction.

TEST INFRASTRUCTURE — see oracle/__init__.py.

The reference does not compute interpolation weights itself: `make-regrid-file`
(commands/make-regrid-file.py:142-160) calls the external `mir` binary through
`earthkit.regrid.utils.mir.mir_make_matrix`, and the `in_grid` / `out_grid` / `method` recipes
(filters/fields/regrid.py:211-259) use earthkit-regrid's pre-computed matrix inventory —
neither is available offline (earthkit-regrid >=0.4,<1, pyproject.toml:41; parity with MIR
unpinned).  What the B200 path builds locally instead is the textbook 4-point bilinear scheme
on a regular lat-lon source grid; this module restates it in numpy (the scipy-built matrix
the device builder `at_bilinear_matrix` is checked against, bit for bit) and is pinned by its
defining properties in tests/test_oracle_matrix.py: rows sum to 1, a linear function of
(lat, lon) is reproduced, a target on a source point takes that point's value.
"""

from __future__ import annotations

import numpy as np
from scipy.sparse import csr_array


def regular_grid_parameters(lat: np.ndarray, lon: np.ndarray):
    """(lat0, dlat, n_lat, lon0, dlon, n_lon) of a row-major regular lat-lon grid given as its
    full point lists, or None when the points are not such a grid."""
    lat, lon = np.asarray(lat, dtype=np.float64).reshape(-1), np.asarray(lon, dtype=np.float64).reshape(-1)
    n = lat.size
    if n < 4 or lon.size != n:
        return None
    n_lon = int(np.argmax(lat != lat[0])) if (lat != lat[0]).any() else 0
    if n_lon < 2 or n % n_lon:
        return None
    n_lat = n // n_lon
    if n_lat < 2:
        return None
    rows, cols = lat.reshape(n_lat, n_lon), lon.reshape(n_lat, n_lon)
    if not (rows == rows[:, :1]).all() or not (cols == cols[:1, :]).all():
        return None
    lat0, lon0 = float(rows[0, 0]), float(cols[0, 0])
    dlat, dlon = float(rows[1, 0] - rows[0, 0]), float(cols[0, 1] - cols[0, 0])
    if dlat == 0.0 or dlon <= 0.0:
        return None
    if np.abs(rows[:, 0] - (lat0 + dlat * np.arange(n_lat))).max() > 1e-6 * abs(dlat):
        return None
    if np.abs(cols[0] - (lon0 + dlon * np.arange(n_lon))).max() > 1e-6 * dlon:
        return None
    if abs(n_lon * dlon - 360.0) > 1e-6 * dlon:  # the builder wraps in longitude
        return None
    return lat0, dlat, n_lat, lon0, dlon, n_lon


def bilinear_matrix(lat0, dlat, n_lat, lon0, dlon, n_lon, tgt_lat, tgt_lon):
    """4-point bilinear weights → (data float32[4n], indices int32[4n], indptr int32[n+1], shape).
    Rows sorted by column (stable), explicit zeros kept."""
    tgt_lat, tgt_lon = np.asarray(tgt_lat, dtype=np.float64).reshape(-1), np.asarray(tgt_lon, dtype=np.float64).reshape(-1)
    n = tgt_lat.shape[0]
    fy = (tgt_lat - lat0) / dlat
    j = np.clip(np.floor(fy).astype(np.int64), 0, n_lat - 2)
    wy = fy - j
    fx = np.mod(tgt_lon - lon0, 360.0) / dlon
    i0 = np.floor(fx).astype(np.int64)
    wx = fx - i0
    i0 %= n_lon
    i1 = (i0 + 1) % n_lon
    cols = np.stack([j * n_lon + i0, j * n_lon + i1, (j + 1) * n_lon + i0, (j + 1) * n_lon + i1], axis=1)
    w = np.stack([(1.0 - wy) * (1.0 - wx), (1.0 - wy) * wx, wy * (1.0 - wx), wy * wx], axis=1)
    order = np.argsort(cols, axis=1, kind="stable")
    cols = np.take_along_axis(cols, order, axis=1)
    w = np.take_along_axis(w, order, axis=1)
    indptr = (4 * np.arange(n + 1)).astype(np.int32)
    return w.astype(np.float32).ravel(), cols.astype(np.int32).ravel(), indptr, (n, n_lat * n_lon)


def bilinear_csr(src_lat, src_lon, tgt_lat, tgt_lon) -> csr_array:
    prm = regular_grid_parameters(src_lat, src_lon)
    if prm is None:
        raise ValueError("source grid is not a regular, longitude-periodic lat-lon grid")
    d, i, p, shape = bilinear_matrix(*prm, tgt_lat, tgt_lon)
    return csr_array((d, i, p), shape=shape)
