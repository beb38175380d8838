"""`regrid` — reference `filters/fields/regrid.py:87-516`.

Same constructor arguments, same interpolator precedence (matrix > mask > method=="nearest"
> earthkit, `_interpolator` 432-467) and the same output wrapping
(`NewLatLonField(NewMetadataField(NewDataField…))`, regrid.py:312).  What changes is where
the arithmetic runs: the per-field Python loop of `_interpolate` (regrid.py:204-208) becomes
one batched device pass —

    MIRMatrix                     CSR staged once in HBM, `at_spmm` over all fields
    ScipyKDTreeNearestNeighbours  `spatial.nearest_grid_points` (device kNN) + `at_gather_rows`
    MaskedRegrid                  `at_gather_rows` with the mask's indices

Host fields are uploaded in chunks and packed point-major; fields already resident in HBM
(outputs of another filter of this package) are used in place.  Outputs stay resident and
are downloaded when `to_numpy()` is first called.
"""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from ... import ekd
from ... import grib
from ...batching import fields_to_batch
from ... import _cabi
from ...device import CsrMatrix, DeviceBatch, StreamedRegrid, gather_rows, require_cuda, results_are_host_bound
from ...fields import new_field_from_device_column, new_field_from_latitudes_longitudes, new_fieldlist_from_list
from ...filter import Filter
from . import filter_registry

LOG = logging.getLogger(__name__)


def _free_device_bytes() -> int:
    """HBM this process can still use, from torch's own counters (cudaMemGetInfo takes tens of
    milliseconds on a 180 GB device — too slow for every forward call)."""
    torch = require_cuda()
    total = torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory
    return int(0.95 * total - torch.cuda.memory_allocated())


def as_gridspec(grid: Any) -> dict[str, Any] | None:
    if grid is None:
        return None
    if isinstance(grid, (str, list, tuple)):
        return {"grid": grid}
    return grid


def as_griddata(grid: Any) -> dict[str, Any] | None:
    """Grid given as a field, a {latitudes, longitudes} dict or a named grid."""
    if grid is None:
        return None
    if isinstance(grid, ekd.Field) or (hasattr(grid, "grid_points") and hasattr(grid, "to_numpy")):
        lat, lon = grid.grid_points()
        return dict(latitudes=lat, longitudes=lon)
    if isinstance(grid, dict) and "latitudes" in grid and "longitudes" in grid:
        return grid
    if isinstance(grid, (str, list, tuple)):
        from ...grids import lookup

        return lookup(grid)
    raise ValueError(f"Invalid grid: {grid}")


@filter_registry.register("regrid")
class RegridFilter(Filter):
    """Regrid fields with a precomputed matrix, a nearest-neighbour search or an index mask."""

    def __init__(
        self,
        *,
        in_grid: Any | None = None,
        out_grid: Any | None = None,
        method: str | None = None,
        matrix: str | None = None,
        mask: str | None = None,
        check: bool = False,
    ) -> None:
        self.in_grid = in_grid
        self.out_grid = out_grid
        self.method = method
        self.interpolator = make_interpolator(in_grid=in_grid, out_grid=out_grid, method=method, matrix=matrix, mask=mask, check=check)

    def forward(self, data: Any) -> Any:
        return self._interpolate(data)

    def _interpolate(self, data: Any) -> Any:
        fields = list(data)
        if not fields:
            return new_fieldlist_from_list([])
        if hasattr(self.interpolator, "regrid_batch"):
            return new_fieldlist_from_list(self.interpolator.regrid_batch(fields))
        return new_fieldlist_from_list([self.interpolator(f) for f in fields])


class _BatchedInterpolator:
    """Shared driver: batch the FieldList, transform on the device, wrap the outputs."""

    def __call__(self, field: Any) -> Any:
        return self.regrid_batch([field])[0]

    #: fraction of the free HBM one sub-batch (inputs + outputs) may take
    memory_fraction = 0.4

    def regrid_batch(self, fields: list[Any]) -> list[Any]:
        """All fields in one device pass.

        Fields that arrive from the host and whose results are headed back to it are *streamed*
        (`at_hostio_regrid`: staging, H2D, compute and D2H of consecutive chunks overlap; only
        the results stay in HBM, and not even those when they would not fit).  Anything else —
        fields already resident, results wanted by the next filter of a pipeline — takes the
        batch path, in sub-batches whose outputs are moved to host memory as soon as they are
        computed when the FieldList does not fit in HBM."""
        from ...fields import device_column_of

        if results_are_host_bound() and all(device_column_of(f) is None for f in fields):
            return self._regrid_streamed(fields)
        per_field = self.bytes_per_field(fields[0])
        limit = max(4, int(self.memory_fraction * _free_device_bytes() // max(1, per_field)) // 4 * 4)
        if len(fields) <= limit:
            return self._regrid_resident(fields)
        LOG.info("regrid: %d fields exceed the device budget (%d per pass): streaming in sub-batches", len(fields), limit)
        out: list[Any] = []
        for i in range(0, len(fields), limit):
            part = self._regrid_resident(fields[i : i + limit])
            for batch in {id(b): b for b, _ in filter(None, (device_column_of(f) for f in part))}.values():
                batch.offload()
            out.extend(part)
        return out

    def bytes_per_field(self, field: Any) -> int:
        """Device bytes one field costs: its input column plus its output column (float64 worst case)."""
        n_in = int(np.prod(field.shape)) if hasattr(field, "shape") else 0
        return 8 * (n_in + self.output_points(n_in))

    def output_points(self, n_in: int) -> int:
        return n_in

    @staticmethod
    def _host_values(field: Any) -> np.ndarray:
        # reshape(-1) instead of flatten=True: no host copy when the field is contiguous
        v = np.asarray(field.to_numpy()).reshape(-1)
        if v.dtype != np.float32 and v.dtype != np.float64:
            v = v.astype(np.float64)
        return v if v.flags.c_contiguous else np.ascontiguousarray(v)

    def _wrap(self, fields: list[Any], idxs: list[int], result: DeviceBatch, out: list[Any]) -> None:
        lat, lon = self.output_grid(fields[idxs[0]])
        for j, i in enumerate(idxs):
            out[i] = new_field_from_latitudes_longitudes(new_field_from_device_column(result, j, template=fields[i]), latitudes=lat, longitudes=lon)

    def _regrid_streamed(self, fields: list[Any]) -> list[Any]:
        torch = require_cuda()
        self.prepare(fields[0])
        out: list[Any] = [None] * len(fields)
        packed, packed_idx, other_idx = grib.split(fields)
        if packed is not None:
            # simple-packed GRIB messages: the packed octets cross PCIe, the device decodes them
            # to float64 (what `to_numpy()` of a GRIB field returns) in front of the transform
            op, csr, index, n_tgt, y_dtype = self.stream_spec(packed.n_points, torch.float32 if packed.dtype == np.float32 else torch.float64)
            keep = 8 * n_tgt * len(packed_idx) <= self.memory_fraction * _free_device_bytes()
            job = StreamedRegrid(op, csr, index, n_tgt, y_dtype, packed, keep_resident=keep, to_host=True)
            try:
                self._wrap(fields, packed_idx, job.batch, out)
            finally:
                job.join()
            if not other_idx:
                return out
        # everything else (numpy fields, other packings, bitmaps): values fetched on the host
        values: dict[int, np.ndarray] = {i: self._host_values(fields[i]) for i in other_idx}
        by_dtype: dict[Any, list[int]] = {}
        for i, v in values.items():
            by_dtype.setdefault(v.dtype, []).append(i)
        for dtype, idxs in by_dtype.items():
            arrays = [values[i] for i in idxs]
            n_src = int(arrays[0].size)
            for i in idxs:
                if values[i].size != n_src:
                    raise ValueError(f"field {i} has {values[i].size} points, expected {n_src}")
            op, csr, index, n_tgt, y_dtype = self.stream_spec(n_src, torch.float32 if dtype == np.float32 else torch.float64)
            keep = 8 * n_tgt * len(arrays) <= self.memory_fraction * _free_device_bytes()
            job = StreamedRegrid(op, csr, index, n_tgt, y_dtype, arrays, keep_resident=keep, to_host=True)
            try:
                self._wrap(fields, idxs, job.batch, out)  # while the fields stream through the GPU
            finally:
                job.join()
        return out

    def stream_spec(self, n_src: int, x_dtype: Any):
        """→ (op, csr, device index, n_tgt, result dtype) of `at_hostio_regrid`."""
        raise NotImplementedError

    def _regrid_resident(self, fields: list[Any]) -> list[Any]:
        self.prepare(fields[0])
        # numpy dtypes differ per field in principle; batch runs of equal dtype together.
        # Host values are fetched once here (to_numpy decodes / copies) and handed to the upload.
        from ...fields import device_column_of

        out: list[Any] = [None] * len(fields)
        if all(device_column_of(f) is None for f in fields):
            packed = grib.packed_of(fields)
            if packed is not None:  # decoded on the device, see _regrid_streamed
                result = self.apply(grib.upload(packed))
                if results_are_host_bound():
                    result.prefetch()
                self._wrap(fields, list(range(len(fields))), result, out)
                return out
        by_dtype: dict[Any, list[int]] = {}
        host_values: dict[int, np.ndarray] = {}
        for i, f in enumerate(fields):
            col = device_column_of(f)
            if col is not None:
                key = str(col[0].data.dtype)
            else:
                host_values[i] = v = self._host_values(f)
                key = "torch.float32" if v.dtype == np.float32 else "torch.float64"
            by_dtype.setdefault(key, []).append(i)
        for _, idxs in by_dtype.items():
            batch = fields_to_batch([fields[i] for i in idxs], host_values=[host_values.get(i) for i in idxs])
            result = self.apply(batch)
            if results_are_host_bound():
                result.prefetch()
            self._wrap(fields, idxs, result, out)
        return out

    def prepare(self, first_field: Any) -> None:
        pass

    def apply(self, batch: DeviceBatch) -> DeviceBatch:
        raise NotImplementedError

    def output_grid(self, field: Any):
        raise NotImplementedError


class MIRMatrix(_BatchedInterpolator):
    """Matrix created by ``anemoi-transform make-regrid-file`` (npz schema
    make-regrid-file.py:150-160), applied as a device SpMM."""

    def __init__(self, *, matrix: str, check: bool) -> None:
        self.check = check
        if self.check:
            LOG.warning("Check is not supported by MIRMatrix")
        loaded = dict(np.load(matrix))
        self.matrix = CsrMatrix(loaded["matrix_data"], loaded["matrix_indices"], loaded["matrix_indptr"], tuple(loaded["matrix_shape"]))
        self.in_grid = dict(latitudes=loaded["in_latitudes"], longitudes=loaded["in_longitudes"])
        self.out_grid = dict(latitudes=loaded["out_latitudes"], longitudes=loaded["out_longitudes"])

    def output_points(self, n_in: int) -> int:
        return self.matrix.shape[0]

    def apply(self, batch: DeviceBatch) -> DeviceBatch:
        if batch.n_points != self.matrix.shape[1]:
            raise ValueError(f"dimension mismatch: matrix has {self.matrix.shape[1]} columns, field has {batch.n_points} points")
        return DeviceBatch(self.matrix.apply(batch.data, n_fields=batch.n_fields), batch.n_fields)

    def stream_spec(self, n_src: int, x_dtype: Any):
        return _cabi.HOSTIO_SPMM, self.matrix, None, self.matrix.shape[0], self.matrix.result_dtype(x_dtype)

    def output_grid(self, field: Any):
        return self.out_grid["latitudes"], self.out_grid["longitudes"]


class ScipyKDTreeNearestNeighbours(_BatchedInterpolator):
    """Nearest-neighbour regridding for grids earthkit-regrid has no matrix for.

    The class name is the reference's (regrid.py:315); the search itself is the bucketed
    device kNN of `spatial.nearest_grid_points`, not scipy.
    """

    nearest_grid_points = None

    def __init__(self, *, in_grid: Any, out_grid: Any, method: str, check: bool = False) -> None:
        if method != "nearest":
            raise NotImplementedError(f"ScipyKDTreeNearestNeighbours does not support {method}, only 'nearest'")
        self.in_grid = as_griddata(in_grid)
        self.out_grid = as_griddata(out_grid)
        if self.out_grid is None:
            raise ValueError("out_grid is required, but not provided")
        if check:
            LOG.warning("Check is not supported by ScipyKDTreeNearestNeighbours")

    def prepare(self, first_field: Any) -> None:
        if self.in_grid is None:
            self.in_grid = as_griddata(first_field)
            assert self.in_grid is not None, first_field
        if self.nearest_grid_points is None:
            from ...spatial import nearest_grid_points

            self.nearest_grid_points = nearest_grid_points(
                self.in_grid["latitudes"],
                self.in_grid["longitudes"],
                self.out_grid["latitudes"],
                self.out_grid["longitudes"],
                _as_device=True,
            )

    def output_points(self, n_in: int) -> int:
        return int(np.size(self.out_grid["latitudes"]))

    def _check_points(self, n_points: int) -> None:
        n_in = (n_points,)
        assert n_in == np.shape(self.in_grid["latitudes"]), (n_in, np.shape(self.in_grid["latitudes"]))
        assert n_in == np.shape(self.in_grid["longitudes"]), (n_in, np.shape(self.in_grid["longitudes"]))

    def apply(self, batch: DeviceBatch) -> DeviceBatch:
        self._check_points(batch.n_points)
        return DeviceBatch(gather_rows(batch.data, self.nearest_grid_points, n_fields=batch.n_fields), batch.n_fields)

    def stream_spec(self, n_src: int, x_dtype: Any):
        self._check_points(n_src)
        idx = self.nearest_grid_points.reshape(-1)
        return _cabi.HOSTIO_GATHER, None, idx, int(idx.shape[0]), x_dtype

    def output_grid(self, field: Any):
        return self.out_grid["latitudes"], self.out_grid["longitudes"]


class MaskedRegrid(_BatchedInterpolator):
    """Select points with a precomputed index (or boolean) mask (`make-regrid-file
    global-on-lam-mask`, make-regrid-file.py:239-240)."""

    out_latitudes = None
    out_longitudes = None

    def __init__(self, *, mask: str, check: bool) -> None:
        if check:
            LOG.warning("Check is not supported by MaskedRegrid")
        self.mask = np.load(mask)["mask"]
        self._device_index = None
        self._index_points = -1

    def _index(self, n_points: int):
        """The mask as device int64 indices into a field of n_points (numpy's indexing rules)."""
        torch = require_cuda()
        if self._device_index is None or self._index_points != n_points:
            m = self.mask
            if m.dtype == np.bool_:
                if m.shape[0] != n_points:
                    raise IndexError(f"boolean index did not match indexed array along axis 0; size of axis is {n_points} but size of corresponding boolean axis is {m.shape[0]}")
                m = np.nonzero(m)[0]
            m = np.asarray(m).astype(np.int64).reshape(-1)
            bad = m[(m < -n_points) | (m >= n_points)]
            if bad.size:
                raise IndexError(f"index {int(bad[0])} is out of bounds for axis 0 with size {n_points}")
            m = np.where(m < 0, m + n_points, m)  # numpy's negative indexing
            self._device_index, self._index_points = torch.from_numpy(m).cuda(), n_points
        return self._device_index

    def apply(self, batch: DeviceBatch) -> DeviceBatch:
        return DeviceBatch(gather_rows(batch.data, self._index(batch.n_points), n_fields=batch.n_fields), batch.n_fields)

    def output_points(self, n_in: int) -> int:
        return int(self._index(n_in).shape[0])

    def stream_spec(self, n_src: int, x_dtype: Any):
        idx = self._index(n_src)
        return _cabi.HOSTIO_GATHER, None, idx, int(idx.shape[0]), x_dtype

    def output_grid(self, field: Any):
        if self.out_latitudes is None or self.out_longitudes is None:
            in_latitudes, in_longitudes = field.grid_points()
            self.out_latitudes = in_latitudes[self.mask]
            self.out_longitudes = in_longitudes[self.mask]
        return self.out_latitudes, self.out_longitudes


class EarthkitRegrid(_BatchedInterpolator):
    """The `in_grid` / `out_grid` / `method` recipes (reference regrid.py:211-259).

    The reference hands these to `earthkit.regrid.interpolate`, i.e. to earthkit-regrid's
    inventory of MIR matrices; neither the package nor the inventory exists offline.  Here the
    matrix is built locally on the device the first time it is needed and applied like any
    other (`MIRMatrix` machinery: one SpMM over the whole FieldList):

        method="linear"                      4-point bilinear weights, regular lat-lon sources
                                             (`regrid_files.make_bilinear_matrix`)
        method="nearest-neighbour" | "nn"    the nearest source point (`make_knn_matrix`, k = 1)

    (`method="nearest"` is dispatched to `ScipyKDTreeNearestNeighbours`, as in the reference.)
    The class name is the reference's so that `_interpolator` / configuration dumps read the
    same; other schemes and non-regular sources raise NotImplementedError and still need a
    `matrix=` made with MIR.
    """

    METHODS = ("linear", "nearest-neighbour", "nn")

    def __init__(self, *, in_grid: Any = None, out_grid: Any = None, method: str = "linear", check: bool = False) -> None:
        if method not in self.METHODS:
            raise NotImplementedError(f"regrid: method {method!r} is not built locally (have {self.METHODS} and 'nearest'); pass `matrix=`")
        if out_grid is None:
            raise ValueError("out_grid is required, but not provided")
        self.in_gridspec = as_gridspec(in_grid)
        self.out_gridspec = as_gridspec(out_grid)
        self.in_grid = as_griddata(in_grid)
        self.out_grid = as_griddata(out_grid)
        self.method = method
        self.matrix: CsrMatrix | None = None
        if check:
            LOG.warning("Check is not supported by EarthkitRegrid")

    def prepare(self, first_field: Any) -> None:
        if self.matrix is not None:
            return
        if self.in_grid is None:  # earthkit-regrid reads the grid off the field; so do we
            self.in_grid = as_griddata(first_field)
        from ...regrid_files import make_bilinear_matrix, make_knn_matrix

        src = (self.in_grid["latitudes"], self.in_grid["longitudes"])
        dst = (self.out_grid["latitudes"], self.out_grid["longitudes"])
        if self.method == "linear":
            data, indices, indptr, shape = make_bilinear_matrix(*src, *dst)
        else:
            data, indices, indptr, shape = make_knn_matrix(*src, *dst, k=1)
        self.matrix = CsrMatrix(data, indices, indptr, shape)

    def output_points(self, n_in: int) -> int:
        return int(np.size(self.out_grid["latitudes"]))

    def apply(self, batch: DeviceBatch) -> DeviceBatch:
        if batch.n_points != self.matrix.shape[1]:
            raise ValueError(f"dimension mismatch: matrix has {self.matrix.shape[1]} columns, field has {batch.n_points} points")
        return DeviceBatch(self.matrix.apply(batch.data, n_fields=batch.n_fields), batch.n_fields)

    def stream_spec(self, n_src: int, x_dtype: Any):
        return _cabi.HOSTIO_SPMM, self.matrix, None, self.matrix.shape[0], self.matrix.result_dtype(x_dtype)

    def output_grid(self, field: Any):
        return self.out_grid["latitudes"], self.out_grid["longitudes"]


def _interpolator(*, method: str | None = None, matrix: str | None = None, mask: str | None = None) -> str:
    if matrix is not None:
        return "MIRMatrix"
    if mask is not None:
        return "MaskedRegrid"
    if method == "nearest":
        return "ScipyKDTreeNearestNeighbours"
    return "EarthkitRegrid"


def make_interpolator(in_grid: Any = None, out_grid: Any = None, method: str | None = None, matrix: str | None = None, mask: str | None = None, check: bool | None = None) -> Any:
    name = _interpolator(method=method, matrix=matrix, mask=mask)
    kwargs = {"in_grid": in_grid, "out_grid": out_grid, "method": method, "matrix": matrix, "mask": mask, "check": check}
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return globals()[name](**kwargs)
