"""The two on-disk artefacts the `regrid` filter consumes — reference
`commands/make-regrid-file.py:142-160` (matrix npz) and `:225-242` (global-on-LAM mask npz).

The file formats and the mask computation live here.  Building an interpolation matrix is
MIR's job in the reference (`earthkit.regrid.utils.mir.mir_make_matrix` + the external `mir`
binary, neither available offline); any scipy sparse matrix can be written with
`save_regrid_matrix`, and two builders run locally on the device:

    make_bilinear_matrix   4-point bilinear weights from a regular, longitude-periodic lat-lon
                           source grid (`at_bilinear_matrix`; bitwise the numpy restatement
                           oracle/matrix.py) — what `regrid(in_grid=…, out_grid=…,
                           method="linear")` uses in place of earthkit-regrid's inventory
    make_knn_matrix        nearest neighbour (k = 1) or inverse-distance weights of the k
                           nearest sources, on the device kNN

Neither is MIR's scheme for reduced Gaussian sources, so fields regridded with them are
comparable with a scipy matrix of the same construction, not with a MIR matrix.
"""

from __future__ import annotations

from typing import Any

import numpy as np


def save_regrid_matrix(output: str, sparse_array: Any, lat1, lon1, lat2, lon2) -> None:
    """Write a CSR matrix with the schema `RegridFilter(matrix=…)` / `MIRMatrix` reads
    (make-regrid-file.py:150-160)."""
    m = sparse_array.tocsr() if hasattr(sparse_array, "tocsr") else sparse_array
    np.savez(
        output,
        matrix_data=m.data,
        matrix_indices=m.indices,
        matrix_indptr=m.indptr,
        matrix_shape=m.shape,
        in_latitudes=lat1,
        in_longitudes=lon1,
        out_latitudes=lat2,
        out_longitudes=lon2,
    )


def make_global_on_lam_mask(lam_lat, lam_lon, global_lat, global_lon, output: str, **kwargs: Any) -> np.ndarray:
    """`anemoi-transform make-regrid-file global-on-lam-mask`: the sorted indices of the global
    points within `distance_km` of a LAM point, computed on the device, written as `mask`
    (what `RegridFilter(mask=…)` / `MaskedRegrid` loads)."""
    from .spatial import global_on_lam_mask

    mask = global_on_lam_mask(lam_lat, lam_lon, global_lat, global_lon, **kwargs)
    np.savez(output, mask=mask)
    return mask


def make_knn_matrix(lat1, lon1, lat2, lon2, output: str | None = None, k: int = 4, power: float = 1.0):
    """A [n_target, n_source] CSR matrix from the k nearest source points of every target
    (device kNN on the unit sphere): weight ∝ 1 / distance**power, rows normalised to 1, columns
    sorted within a row; a target that coincides with a source takes that source alone.
    k = 1 is nearest-neighbour regridding as a matrix.  Written with the regrid-file schema when
    `output` is given.  → (data float32, indices int32, indptr int32, shape)."""
    from .device import KnnIndex
    from .spatial import latlon_to_xyz

    lat1, lon1, lat2, lon2 = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (lat1, lon1, lat2, lon2))
    n_src, n_tgt = lat1.size, lat2.size
    if not 1 <= k <= n_src:
        raise ValueError(f"k={k} must be between 1 and the number of source points ({n_src})")
    index = KnnIndex(latlon_to_xyz(lat1, lon1))
    try:
        idx, dist, _ = index.query(latlon_to_xyz(lat2, lon2), k=k)
        idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    finally:
        index.close()
    exact = dist[:, 0] == 0.0
    with np.errstate(divide="ignore"):
        w = 1.0 / dist**power
    w[exact] = 0.0
    w[exact, 0] = 1.0
    w /= w.sum(axis=1, keepdims=True)
    order = np.argsort(idx, axis=1, kind="stable")
    indices = np.take_along_axis(idx, order, axis=1).astype(np.int32).ravel()
    data = np.take_along_axis(w, order, axis=1).astype(np.float32).ravel()
    indptr = (k * np.arange(n_tgt + 1)).astype(np.int32)
    shape = (n_tgt, n_src)
    if output is not None:
        np.savez(output, matrix_data=data, matrix_indices=indices, matrix_indptr=indptr, matrix_shape=np.asarray(shape),
                 in_latitudes=lat1, in_longitudes=lon1, out_latitudes=lat2, out_longitudes=lon2)  # fmt: skip
    return data, indices, indptr, shape


def regular_grid_parameters(lat, lon):
    """(lat0, dlat, n_lat, lon0, dlon, n_lon) when (lat, lon) are the point lists of a row-major
    regular lat-lon grid that wraps in longitude, else None."""
    lat, lon = np.asarray(lat, dtype=np.float64).reshape(-1), np.asarray(lon, dtype=np.float64).reshape(-1)
    n = lat.size
    if n < 4 or lon.size != n:
        return None
    change = np.flatnonzero(lat != lat[0])
    n_lon = int(change[0]) if change.size else 0
    if n_lon < 2 or n % n_lon or n // n_lon < 2:
        return None
    n_lat = n // n_lon
    rows, cols = lat.reshape(n_lat, n_lon), lon.reshape(n_lat, n_lon)
    if not (rows == rows[:, :1]).all() or not (cols == cols[:1, :]).all():
        return None
    lat0, lon0 = float(rows[0, 0]), float(cols[0, 0])
    dlat, dlon = float(rows[1, 0] - rows[0, 0]), float(cols[0, 1] - cols[0, 0])
    if dlat == 0.0 or dlon <= 0.0:
        return None
    regular = (
        np.abs(rows[:, 0] - (lat0 + dlat * np.arange(n_lat))).max() <= 1e-6 * abs(dlat)
        and np.abs(cols[0] - (lon0 + dlon * np.arange(n_lon))).max() <= 1e-6 * dlon
        and abs(n_lon * dlon - 360.0) <= 1e-6 * dlon
    )
    return (lat0, dlat, n_lat, lon0, dlon, n_lon) if regular else None


def make_bilinear_matrix(lat1, lon1, lat2, lon2, output: str | None = None):
    """A [n_target, n_source] CSR matrix of 4-point bilinear weights from a regular,
    longitude-periodic lat-lon source grid, built on the device.  Four entries per row, sorted
    by column, explicit zeros kept (they propagate NaN like scipy).  Written with the
    regrid-file schema when `output` is given.  → (data float32, indices int32, indptr int32, shape)."""
    from ctypes import c_void_p

    from ._cabi import call
    from .device import require_cuda, stream_ptr, to_device_f64

    torch = require_cuda()
    prm = regular_grid_parameters(lat1, lon1)
    if prm is None:
        raise NotImplementedError(
            "linear interpolation is built locally for regular, longitude-periodic lat-lon source grids only; "
            "pass `matrix=` (make-regrid-file) for other sources, or method='nearest'"
        )
    lat2, lon2 = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (lat2, lon2))
    n_tgt = int(lat2.size)
    lo, hi = sorted((prm[0], prm[0] + prm[1] * (prm[2] - 1)))
    if n_tgt and (lat2.min() < lo - 1e-9 or lat2.max() > hi + 1e-9):
        raise ValueError(f"target latitudes [{lat2.min()}, {lat2.max()}] leave the source grid's [{lo}, {hi}]")
    tlat, tlon = to_device_f64(lat2), to_device_f64(lon2)
    data = torch.empty((4 * n_tgt,), dtype=torch.float32, device="cuda")
    indices = torch.empty((4 * n_tgt,), dtype=torch.int32, device="cuda")
    call("at_bilinear_matrix", *prm[:3], *prm[3:], c_void_p(tlat.data_ptr()), c_void_p(tlon.data_ptr()), n_tgt, c_void_p(data.data_ptr()), c_void_p(indices.data_ptr()), stream_ptr())
    data, indices = data.cpu().numpy(), indices.cpu().numpy()
    indptr = (4 * np.arange(n_tgt + 1)).astype(np.int32)
    shape = (n_tgt, prm[2] * prm[5])
    if output is not None:
        np.savez(output, matrix_data=data, matrix_indices=indices, matrix_indptr=indptr, matrix_shape=np.asarray(shape),
                 in_latitudes=np.asarray(lat1), in_longitudes=np.asarray(lon1), out_latitudes=lat2, out_longitudes=lon2)  # fmt: skip
    return data, indices, indptr, shape
