"""The two on-disk artefacts the `regrid` filter consumes — reference
`commands/make-regrid-file.py:142-160` (matrix npz) and `:225-242` (global-on-LAM mask npz).

The file formats and the mask computation live here.  Building an interpolation matrix is
MIR's job in the reference (`earthkit.regrid.utils.mir.mir_make_matrix` + the external `mir`
binary, neither available offline); any scipy sparse matrix can be written with
`save_regrid_matrix`, and `make_knn_matrix` builds one locally on the device kNN (nearest
neighbour for k = 1, inverse-distance weights of the k nearest sources otherwise — NOT MIR's
schemes, so its fields are not comparable with a MIR matrix's, only with themselves).
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
