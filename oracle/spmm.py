"""Oracle: the regrid matrix application.

TEST INFRASTRUCTURE — see oracle/__init__.py.

Reference: `MIRMatrix.__call__`, filters/fields/regrid.py:309-310 —
    data = field.to_numpy(flatten=True);  data = self.matrix @ data
with `self.matrix = scipy.sparse.csr_array((data, indices, indptr), shape)` (regrid.py:283-285).
scipy is a transitive dependency of the reference that IS installed here and on the GPU
box, so `mir_matrix_apply` below simply makes the reference's call — it is the reference's
arithmetic, not a re-derivation.

`csr_matvec_sequential` restates what scipy's C++ `csr_matvec` does (sequential, unfused
accumulation in storage order from 0) in numpy, and `csr_matvec.c` restates it in plain C
(the multi-threaded CPU baseline of bench.py).  PINNED: both are checked bit for bit against
scipy in tests/test_oracle_spmm.py, and the scipy call is checked against the imported
reference's `MIRMatrix.__call__` (tests/golden/regrid_*.npz, oracle/make_golden.py).
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
from scipy.sparse import csr_array

HERE = Path(__file__).resolve().parent
C_SOURCE = HERE / "csr_matvec.c"
C_LIBRARY = HERE / "_build" / "liboracle_csr.so"


def load_matrix(npz_path) -> tuple[csr_array, dict, dict]:
    """regrid.py:281-290: npz → (csr_array, in_grid, out_grid)."""
    loaded = dict(np.load(npz_path))
    m = csr_array((loaded["matrix_data"], loaded["matrix_indices"], loaded["matrix_indptr"]), shape=loaded["matrix_shape"])
    return (
        m,
        dict(latitudes=loaded["in_latitudes"], longitudes=loaded["in_longitudes"]),
        dict(latitudes=loaded["out_latitudes"], longitudes=loaded["out_longitudes"]),
    )


def mir_matrix_apply(matrix: csr_array, field_values: np.ndarray) -> np.ndarray:
    """regrid.py:309-310 for one field."""
    return matrix @ field_values


def regrid_fields(matrix: csr_array, fields: list[np.ndarray]) -> list[np.ndarray]:
    """The per-field loop of RegridFilter._interpolate, regrid.py:204-208."""
    return [mir_matrix_apply(matrix, np.asarray(f).reshape(-1)) for f in fields]


def csr_matvec_sequential(indptr, indices, data, x) -> np.ndarray:
    """y[i] = (((0 + a0·x0) + a1·x1) + …) in storage order, multiply and add rounded
    separately, in result_type(data, x) — scipy's csr_matvec.  Vectorised over rows by
    position-in-row; numpy's `*` and `+` are separate ufunc calls, so nothing is fused."""
    indptr = np.asarray(indptr).astype(np.int64)
    dtype = np.result_type(data.dtype, x.dtype)
    a = data.astype(dtype, copy=False)
    xv = x.astype(dtype, copy=False)
    n = indptr.shape[0] - 1
    lengths = np.diff(indptr)
    y = np.zeros(n, dtype=dtype)
    with np.errstate(invalid="ignore", over="ignore"):
        for j in range(int(lengths.max()) if n else 0):
            rows = np.nonzero(lengths > j)[0]
            p = indptr[rows] + j
            y[rows] = y[rows] + a[p] * xv[indices[p]]
    return y


# ---- plain-C port (multi-threaded CPU baseline) ----------------------------------------
def build_c_library(force: bool = False) -> Path:
    C_LIBRARY.parent.mkdir(exist_ok=True)
    if force or not C_LIBRARY.exists() or C_LIBRARY.stat().st_mtime < C_SOURCE.stat().st_mtime:
        cmd = ["gcc", "-O3", "-march=x86-64-v2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", str(C_SOURCE), "-o", str(C_LIBRARY)]
        subprocess.run(cmd, check=True)
    return C_LIBRARY


_clib = None


def _c():
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(str(build_c_library()))
        lib.oracle_csr_matvecs_f32.restype = None
        lib.oracle_csr_matvecs_f32.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        lib.oracle_max_threads.restype = ctypes.c_int
        _clib = lib
    return _clib


def c_max_threads() -> int:
    return int(_c().oracle_max_threads())


def c_regrid_fields_f32(indptr, indices, data, fields_in: np.ndarray, n_threads: int = 0) -> np.ndarray:
    """Field-major float32 [F, n_src] → [F, n_tgt]: one csr_matvec per field, fields spread
    over OpenMP threads — the reference's per-field loop on all host cores."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float32)
    fields_in = np.ascontiguousarray(fields_in, dtype=np.float32)
    n_fields, n_src = fields_in.shape
    n_tgt = indptr.shape[0] - 1
    out = np.empty((n_fields, n_tgt), dtype=np.float32)
    _c().oracle_csr_matvecs_f32(
        n_tgt, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, fields_in.ctypes.data, out.ctypes.data, n_fields, n_src, n_tgt, int(n_threads)
    )
    return out
