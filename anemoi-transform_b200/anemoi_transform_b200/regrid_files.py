"""The two on-disk artefacts the `regrid` filter consumes — reference
`commands/make-regrid-file.py:142-160` (matrix npz) and `:225-242` (global-on-LAM mask npz).

Only the file formats and the mask computation live here.  Building an interpolation matrix
is MIR's job in the reference (`earthkit.regrid.utils.mir.mir_make_matrix` + the external
`mir` binary); any scipy sparse matrix can be written with `save_regrid_matrix`.
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
