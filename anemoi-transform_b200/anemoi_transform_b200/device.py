"""Device-side objects of the hot path: thin Python owners of the C-ABI handles.

torch is used for device memory, streams and (elsewhere) torch.distributed only; every
kernel launched from here lives in libat_b200.so.
"""

from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_int64, c_void_p
from typing import Iterable, Sequence

import numpy as np

from . import _cabi
from ._cabi import AT_F32, AT_F64, AT_I32, AT_I64, EpiCol, EpiSegment, call


def _torch():
    import torch

    return torch


def require_cuda():
    """Fail loudly when the CUDA path cannot run (no silent CPU fallback)."""
    torch = _torch()
    _cabi.load()
    if not torch.cuda.is_available():
        raise _cabi.NativeLibraryError("no CUDA device available: anemoi_transform_b200 has no CPU fallback")
    return torch


def stream_ptr() -> c_void_p:
    torch = _torch()
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t) -> c_void_p:
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(None)


def _np_dtype_code(a: np.ndarray, kinds: dict) -> int:
    try:
        return kinds[a.dtype.type]
    except KeyError:
        raise TypeError(f"unsupported dtype {a.dtype}") from None


_FLOAT_CODES = {np.float32: AT_F32, np.float64: AT_F64}
_INT_CODES = {np.int32: AT_I32, np.int64: AT_I64}


def round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def empty_batch(n_rows: int, n_cols: int, dtype, device, written: int | None = None):
    """[n_rows, round_up(n_cols, 4)] without a full memset: the kernels write every real column;
    only the (at most 3) padding columns, or whatever lies beyond `written`, are zeroed."""
    torch = _torch()
    ld = round_up(n_cols, 4)
    out = torch.empty((n_rows, ld), dtype=dtype, device=device)
    first_unwritten = n_cols if written is None else written
    if first_unwritten < ld:
        out[:, first_unwritten:].zero_()
    return out


class CsrMatrix:
    """A CSR interpolation matrix staged once in HBM (reference: MIRMatrix.__init__,
    filters/fields/regrid.py:281-285)."""

    def __init__(self, data: np.ndarray, indices: np.ndarray, indptr: np.ndarray, shape: Sequence[int]):
        require_cuda()
        data = np.ascontiguousarray(data)
        indices = np.ascontiguousarray(indices)
        indptr = np.ascontiguousarray(indptr)
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        if indices.dtype not in (np.int32, np.int64):
            indices = indices.astype(np.int64)
        if indptr.dtype not in (np.int32, np.int64):
            indptr = indptr.astype(np.int64)
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = int(data.shape[0])
        self.dtype = data.dtype
        if indptr.shape[0] != self.shape[0] + 1:
            raise ValueError(f"indptr has {indptr.shape[0]} entries for {self.shape[0]} rows")
        if indices.shape[0] != self.nnz:
            raise ValueError("indices and data differ in length")
        h = c_void_p()
        call(
            "at_csr_create",
            self.shape[0],
            self.shape[1],
            self.nnz,
            indptr.ctypes.data_as(c_void_p),
            _np_dtype_code(indptr, _INT_CODES),
            indices.ctypes.data_as(c_void_p),
            _np_dtype_code(indices, _INT_CODES),
            data.ctypes.data_as(c_void_p),
            _np_dtype_code(data, _FLOAT_CODES),
            byref(h),
        )
        self._h = h
        u = c_int()
        call("at_csr_info", self._h, None, None, None, byref(u), None)
        self.uniform_nnz = u.value

    @classmethod
    def from_scipy(cls, m) -> "CsrMatrix":
        m = m.tocsr()
        return cls(m.data, m.indices, m.indptr, m.shape)

    @property
    def handle(self) -> c_void_p:
        if self._h is None:
            raise RuntimeError("CsrMatrix used after close()")
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            call("at_csr_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def result_dtype(self, x_dtype):
        torch = _torch()
        if self.dtype == np.float64 or x_dtype == torch.float64:
            return torch.float64
        return torch.float32

    def apply(self, X, n_fields: int | None = None, out=None, variant: int = 0):
        """Y[n_rows, ld] = A · X[n_cols, ld]  (point-major batches, torch CUDA tensors)."""
        torch = _torch()
        if X.dim() != 2 or X.shape[0] != self.shape[1]:
            raise ValueError(f"X has shape {tuple(X.shape)}, expected [{self.shape[1]}, fields]")
        if X.stride(1) != 1:
            raise ValueError("X must be row-major (unit stride along fields)")
        n_fields = X.shape[1] if n_fields is None else n_fields
        ydt = self.result_dtype(X.dtype)
        if out is None:
            out = torch.empty((self.shape[0], X.shape[1]), dtype=ydt, device=X.device)
        code = {torch.float32: AT_F32, torch.float64: AT_F64}
        call(
            "at_spmm",
            self.handle,
            _ptr(X),
            code[X.dtype],
            X.stride(0),
            _ptr(out),
            code[out.dtype],
            out.stride(0),
            n_fields,
            variant,
            stream_ptr(),
        )
        return out


class Epilogue:
    """A fused pointwise program (segments + per-output-column parameters)."""

    def __init__(self, segments: Iterable[tuple], cols: Iterable[tuple[float, float, float, int]]):
        """segments: (kind, in_col, n_in, out_col[, pa, pb]); cols: (lo, hi, pressure, flags) per output column."""
        require_cuda()
        segs = [tuple(s) + (0.0, 0.0)[len(s) - 4 :] for s in segments]
        cols = list(cols)
        seg_arr = (EpiSegment * len(segs))(*[EpiSegment(int(s[0]), int(s[1]), int(s[2]), int(s[3]), float(s[4]), float(s[5])) for s in segs])
        col_arr = (EpiCol * len(cols))(*[EpiCol(float(lo), float(hi), float(p), int(fl), 0) for lo, hi, p, fl in cols])
        h = c_void_p()
        call("at_epilogue_create", seg_arr, len(segs), col_arr, len(cols), byref(h))
        self._h = h
        self.n_out_cols = len(cols)
        self.n_in_cols = max(s[1] + s[2] for s in segs)

    @property
    def handle(self) -> c_void_p:
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            call("at_epilogue_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, X, out=None, row_mask=None):
        """Y = epilogue(X) on a resident point-major batch."""
        torch = _torch()
        if out is None:
            out = empty_batch(X.shape[0], self.n_out_cols, X.dtype, X.device)
        code = {torch.float32: AT_F32, torch.float64: AT_F64}[X.dtype]
        call("at_pointwise", self.handle, X.shape[0], _ptr(X), X.stride(0), _ptr(out), out.stride(0), code, _ptr(row_mask), stream_ptr())
        return out

    def apply_fused(self, csr: CsrMatrix, X, out=None, row_mask=None):
        """Y = epilogue(A · X)."""
        torch = _torch()
        if out is None:
            out = empty_batch(csr.shape[0], self.n_out_cols, torch.float32, X.device)
        call("at_spmm_fused", csr.handle, self.handle, _ptr(X), X.stride(0), _ptr(out), out.stride(0), _ptr(row_mask), stream_ptr())
        return out


class KnnIndex:
    """Bucketed search structure over float64 xyz sources (replaces cKDTree(points))."""

    def __init__(self, xyz, cell_size: float = 0.0):
        """xyz: tuple of three float64 arrays (numpy on host or torch CUDA tensors)."""
        torch = require_cuda()
        x, y, z = xyz
        if isinstance(x, torch.Tensor):
            x, y, z = (t.contiguous().to(torch.float64) for t in (x, y, z))
            self.n = int(x.shape[0])
            ptrs = (_ptr(x), _ptr(y), _ptr(z))
            on_device = 1
            torch.cuda.current_stream().synchronize()
        else:
            x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
            self.n = int(x.shape[0])
            ptrs = tuple(a.ctypes.data_as(c_void_p) for a in (x, y, z))
            on_device = 0
        h = c_void_p()
        call("at_knn_create", *ptrs, self.n, on_device, float(cell_size), byref(h))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            call("at_knn_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, qxyz, k: int = 1, distance_upper_bound: float = float("inf"), want_ties: bool = False):
        """→ (idx int64[nq,k], dist float64[nq,k], ties uint8[nq] | None), CUDA tensors."""
        torch = _torch()
        qx, qy, qz = (to_device_f64(a) for a in qxyz)
        nq = int(qx.shape[0])
        idx = torch.empty((nq, k), dtype=torch.int64, device=qx.device)
        dist = torch.empty((nq, k), dtype=torch.float64, device=qx.device)
        ties = torch.empty((nq,), dtype=torch.uint8, device=qx.device) if want_ties else None
        call("at_knn_query", self._h, _ptr(qx), _ptr(qy), _ptr(qz), nq, int(k), float(distance_upper_bound), _ptr(idx), _ptr(dist), _ptr(ties), stream_ptr())
        return idx, dist, ties

    def ball_mark(self, qxyz, r: float, mark=None):
        """mark[j] |= any query within r of source j.  → uint8[n] CUDA tensor."""
        torch = _torch()
        qx, qy, qz = (to_device_f64(a) for a in qxyz)
        if mark is None:
            mark = torch.zeros((self.n,), dtype=torch.uint8, device=qx.device)
        call("at_ball_mark", self._h, _ptr(qx), _ptr(qy), _ptr(qz), int(qx.shape[0]), float(r), _ptr(mark), stream_ptr())
        return mark

    def min_nn_distance(self, first: int = 0, count: int = -1) -> float:
        """min over sources [first, first+count) of the distance to their 2nd nearest source."""
        out = c_double()
        call("at_min_nn_distance", self._h, int(first), int(count), byref(out), stream_ptr())
        return out.value


def to_device_f64(a):
    torch = _torch()
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def compact_mask(mark):
    """Sorted int64 indices of the non-zero bytes of a uint8 CUDA tensor."""
    torch = _torch()
    n = int(mark.shape[0])
    out = torch.empty((max(n, 1),), dtype=torch.int64, device=mark.device)
    count = c_int64()
    call("at_compact_mask", _ptr(mark), n, _ptr(out), byref(count), stream_ptr())
    return out[: count.value]


def cropping_mask_device(lats, lons, north, west, south, east):
    torch = _torch()
    la, lo = to_device_f64(lats), to_device_f64(lons)
    mask = torch.empty((la.shape[0],), dtype=torch.uint8, device=la.device)
    call("at_cropping_mask", _ptr(la), _ptr(lo), int(la.shape[0]), float(north), float(west), float(south), float(east), _ptr(mask), stream_ptr())
    return mask


def transpose(src, out=None):
    """out[c, r] = src[r, c] for 2-D CUDA tensors of 4- or 8-byte elements."""
    torch = _torch()
    rows, cols = src.shape
    if src.stride(1) != 1:
        raise ValueError("transpose: source must have unit stride along its last axis")
    if out is None:
        out = torch.empty((cols, rows), dtype=src.dtype, device=src.device)
    call("at_transpose", _ptr(src), rows, cols, src.stride(0), _ptr(out), out.stride(0), src.element_size(), stream_ptr())
    return out


def gather_rows(X, idx, n_fields: int | None = None, out=None):
    """out[i, :] = X[idx[i], :] (numpy's data[..., idx] on a point-major batch)."""
    torch = _torch()
    idx = idx.to(device=X.device, dtype=torch.int64).contiguous()
    n_out = int(idx.shape[0])
    n_fields = X.shape[1] if n_fields is None else n_fields
    if out is None:
        # whole 16-byte chunks are copied, padding columns included, when the layout allows it
        per16 = 16 // X.element_size()
        vec16 = X.stride(0) % per16 == 0
        out = empty_batch(n_out, X.shape[1], X.dtype, X.device, written=min(X.shape[1], round_up(n_fields, per16)) if vec16 else n_fields)
    err = torch.zeros((1,), dtype=torch.int32, device=X.device)
    call("at_gather_rows", _ptr(idx), n_out, X.shape[0], _ptr(X), X.stride(0), _ptr(out), out.stride(0), n_fields, X.element_size(), _ptr(err), stream_ptr())
    if int(err.item()) != 0:
        raise IndexError(f"index out of bounds for axis with size {X.shape[0]}")
    return out


def gather_cols(X, cols: Sequence[int], out=None):
    """out[:, j] = X[:, cols[j]] — regroup the fields of a resident point-major batch."""
    torch = _torch()
    n_out = len(cols)
    if any(c < 0 or c >= X.shape[1] for c in cols):
        raise IndexError("column index out of range")
    if out is None:
        out = empty_batch(X.shape[0], n_out, X.dtype, X.device)
    index = torch.tensor(list(cols), dtype=torch.int32, device=X.device)
    call("at_gather_cols", _ptr(index), n_out, X.shape[0], _ptr(X), X.stride(0), _ptr(out), out.stride(0), X.element_size(), stream_ptr())
    return out


def sum_cols(X, cols: Sequence[int], n_groups: int, n_terms: int, out=None):
    """out[:, g] = X[:, cols[g*n_terms]] + X[:, cols[g*n_terms+1]] + … (sequential, X's dtype)."""
    torch = _torch()
    if len(cols) != n_groups * n_terms or any(c < 0 or c >= X.shape[1] for c in cols):
        raise IndexError("sum_cols: bad column list")
    if out is None:
        out = empty_batch(X.shape[0], n_groups, X.dtype, X.device)
    index = torch.tensor(list(cols), dtype=torch.int32, device=X.device)
    code = AT_F32 if X.dtype == torch.float32 else AT_F64
    call("at_sum_cols", _ptr(index), n_groups, n_terms, X.shape[0], _ptr(X), X.stride(0), _ptr(out), out.stride(0), code, stream_ptr())
    return out


def range_flags(X, first_col: int, n_cols: int, lo: float, hi: float) -> list[int]:
    """Per column: bit 0 any value < lo, bit 1 any value > hi, bit 2 any NaN."""
    torch = _torch()
    flags = torch.zeros((max(n_cols, 1),), dtype=torch.int32, device=X.device)
    code = AT_F32 if X.dtype == torch.float32 else AT_F64
    call("at_range_flags", _ptr(X), X.stride(0), X.shape[0], int(first_col), int(n_cols), code, float(lo), float(hi), _ptr(flags), stream_ptr())
    return [int(v) for v in flags[:n_cols].cpu().tolist()]


def compare_mask(values, op: int, threshold: float):
    """uint8 mask = OP(values, threshold) for a 1-D (possibly strided) float CUDA tensor."""
    torch = _torch()
    if values.dtype not in (torch.float32, torch.float64):
        values = values.to(torch.float64)
    n = int(values.shape[0])
    mask = torch.empty((n,), dtype=torch.uint8, device=values.device)
    code = AT_F32 if values.dtype == torch.float32 else AT_F64
    call("at_compare_mask", _ptr(values), code, values.stride(0) if n > 1 else 1, n, int(op), float(threshold), _ptr(mask), stream_ptr())
    return mask


_STAGING_BYTES = 128 << 20  # per slot; two slots pinned + two on the device, allocated once
_COPY_THREADS = None
_N_COPY_THREADS = 1
_RING = None


def _chunk_fields(n_fields: int, field_bytes: int, chunk: int | None) -> int:
    if chunk is None:
        chunk = max(1, min(256, _STAGING_BYTES // max(1, field_bytes)))
    return max(1, min(int(chunk), n_fields))


def _parallel_copy(pairs) -> None:
    """dst[...] = src for (dst, src) numpy pairs; big batches are spread over host threads
    (numpy releases the GIL while it copies).  One task per thread, not per field: a pool task
    costs tens of microseconds, a 260 KB field copies in about as long."""
    global _COPY_THREADS, _N_COPY_THREADS
    total = sum(d.nbytes for d, _ in pairs)
    if total < (4 << 20):
        for d, s in pairs:
            np.copyto(d, s, casting="unsafe")
        return
    if _COPY_THREADS is None:
        import os
        from concurrent.futures import ThreadPoolExecutor

        _N_COPY_THREADS = int(os.environ.get("AT_B200_COPY_THREADS", 0)) or max(1, min(8, (len(os.sched_getaffinity(0)) or 2) - 1))
        _COPY_THREADS = ThreadPoolExecutor(max_workers=_N_COPY_THREADS, thread_name_prefix="at-copy")
    if len(pairs) < _N_COPY_THREADS:  # few large fields: split them
        parts = -(-_N_COPY_THREADS // len(pairs))
        split = []
        for d, s in pairs:
            cuts = np.linspace(0, d.size, parts + 1).astype(np.int64)
            split += [(d[a:b], s[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        pairs = split
    n_groups = min(_N_COPY_THREADS, len(pairs))
    groups = [pairs[g::n_groups] for g in range(n_groups)]

    def work(group):
        for d, s in group:
            np.copyto(d, s, casting="unsafe")

    futures = [_COPY_THREADS.submit(work, g) for g in groups[1:]]
    work(groups[0])  # the calling thread takes a share
    for f in futures:
        f.result()


class _StagingRing:
    """Two pinned host slots and two device slots of raw bytes, reused by every upload /
    download of the process (pinning memory is slow; do it once)."""

    def __init__(self, nbytes: int):
        torch = _torch()
        self.nbytes = nbytes
        self.pinned = [torch.empty((nbytes,), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.device = [torch.empty((nbytes,), dtype=torch.uint8, device="cuda") for _ in range(2)]
        self.events = [None, None]
        self.busy = False

    def wait(self, slot: int) -> None:
        if self.events[slot] is not None:
            self.events[slot].synchronize()
            self.events[slot] = None

    def record(self, slot: int) -> None:
        torch = _torch()
        ev = torch.cuda.Event()
        ev.record()
        self.events[slot] = ev

    def release(self) -> None:
        self.wait(0)
        self.wait(1)
        self.busy = False


def _staging_ring(nbytes: int) -> _StagingRing:
    """The process-wide ring, grown when a chunk needs more (per device, single-threaded use)."""
    global _RING
    torch = _torch()
    nbytes = round_up(max(nbytes, 1 << 20), 1 << 20)
    dev = torch.cuda.current_device()
    if _RING is None or _RING.nbytes < nbytes or _RING.device[0].device.index != dev or _RING.busy:
        ring = _StagingRing(nbytes)
        if _RING is None or not _RING.busy:
            _RING = ring
    else:
        ring = _RING
    ring.busy = True
    return ring


class DeviceBatch:
    """A batch of fields resident in HBM, point-major: data[n_points, ld], ld % 4 == 0."""

    def __init__(self, data, n_fields: int):
        self.data = data
        self.n_fields = int(n_fields)
        self._host = None  # per-column host arrays downloaded at the first to_numpy(), handed out once

    @property
    def n_points(self) -> int:
        return int(self.data.shape[0]) if self.data is not None else self._n_points

    @property
    def resident(self) -> bool:
        return self.data is not None

    @classmethod
    def from_host_fields(cls, arrays: Sequence[np.ndarray], chunk: int | None = None) -> "DeviceBatch":
        """Upload F host fields (each [n_points]) and pack them point-major.

        Fields go up in chunks: host threads copy a chunk's arrays from wherever they live
        (pageable memory as a rule) into a pinned staging slot, one async H2D moves the slot
        into a field-major device buffer, `at_transpose` writes it into its columns of the
        batch.  Two slots, so the host copies of chunk k+1 overlap the DMA of chunk k."""
        torch = require_cuda()
        n_fields = len(arrays)
        if n_fields == 0:
            raise ValueError("empty batch")
        views = [np.asarray(a).reshape(-1) for a in arrays]
        n_points = int(views[0].size)
        dtype = np.result_type(*{v.dtype for v in views})
        if dtype not in (np.float32, np.float64):
            dtype = np.dtype(np.float64)
        tdtype = torch.float32 if dtype == np.float32 else torch.float64
        for i, v in enumerate(views):
            if v.size != n_points:
                raise ValueError(f"field {i} has {v.size} points, expected {n_points}")
        ld = round_up(n_fields, 4)
        pm = empty_batch(n_points, n_fields, tdtype, "cuda")
        chunk = _chunk_fields(n_fields, n_points * dtype.itemsize, chunk)
        ring = _staging_ring(chunk * n_points * dtype.itemsize)
        pins = [r.view(tdtype)[: chunk * n_points].view(chunk, n_points) for r in ring.pinned]
        devs = [r.view(tdtype)[: chunk * n_points].view(chunk, n_points) for r in ring.device]
        try:
            for k, c0 in enumerate(range(0, n_fields, chunk)):
                slot = k % 2
                nf = min(chunk, n_fields - c0)
                ring.wait(slot)  # the DMA that last read this pinned slot
                _parallel_copy([(pins[slot][j].numpy(), views[c0 + j]) for j in range(nf)])
                devs[slot][:nf].copy_(pins[slot][:nf], non_blocking=True)
                ring.record(slot)
                call("at_transpose", _ptr(devs[slot]), nf, n_points, n_points, c_void_p(pm.data_ptr() + c0 * pm.element_size()), ld, pm.element_size(), stream_ptr())
        finally:
            ring.release()
        return cls(pm, n_fields)

    def _download(self, dests: Sequence[np.ndarray], first_col: int = 0, chunk: int | None = None) -> None:
        """dests[j][:] = column first_col + j: chunks are transposed field-major on the device,
        moved into a pinned slot by one async D2H, and copied into the destinations by host
        threads while the next chunk is in flight."""
        torch = _torch()
        tdtype = self.data.dtype
        n_points, esz, n = self.n_points, self.data.element_size(), len(dests)
        chunk = _chunk_fields(n, n_points * esz, chunk)
        ring = _staging_ring(chunk * n_points * esz)
        pins = [r.view(tdtype)[: chunk * n_points].view(chunk, n_points) for r in ring.pinned]
        devs = [r.view(tdtype)[: chunk * n_points].view(chunk, n_points) for r in ring.device]
        pending = None

        def drain(p):
            slot, c0, nf = p
            ring.wait(slot)
            _parallel_copy([(dests[c0 + j], pins[slot][j].numpy()) for j in range(nf)])

        try:
            for k, c0 in enumerate(range(0, n, chunk)):
                slot = k % 2
                nf = min(chunk, n - c0)
                call("at_transpose", c_void_p(self.data.data_ptr() + (first_col + c0) * esz), n_points, nf, self.data.stride(0), _ptr(devs[slot]), n_points, esz, stream_ptr())
                pins[slot][:nf].copy_(devs[slot][:nf], non_blocking=True)
                ring.record(slot)
                if pending is not None:
                    drain(pending)
                pending = (slot, c0, nf)
            if pending is not None:
                drain(pending)
        finally:
            ring.release()

    def _np_dtype(self):
        if self.data is None:
            return self._host[0].dtype
        return np.dtype(np.float32 if self.data.dtype == _torch().float32 else np.float64)

    def to_host_fields(self, chunk: int | None = None) -> np.ndarray:
        """→ numpy [n_fields, n_points] (field-major), a fresh array."""
        host = np.empty((self.n_fields, self.n_points), dtype=self._np_dtype())
        self._download([host[j] for j in range(self.n_fields)], 0, chunk)
        return host

    def offload(self) -> None:
        """Move the batch to host memory and free its HBM (used when a FieldList is larger than
        the device: outputs of finished sub-batches must not pile up in HBM).  From then on
        `take_column` hands out copies of the host arrays."""
        if self.data is None:
            return
        if self._host is None:
            self._host = [np.empty((self.n_points,), dtype=self._np_dtype()) for _ in range(self.n_fields)]
            self._download(self._host)
        else:  # columns already handed out were the callers'; fetch them again for our own copy
            missing = [j for j, a in enumerate(self._host) if a is None]
            for j in missing:
                self._host[j] = np.empty((self.n_points,), dtype=self._np_dtype())
                self._download([self._host[j]], first_col=j)
        self._n_points = self.n_points
        self.data = None

    def take_column(self, col: int) -> np.ndarray:
        """A fresh host array with the values of column `col`, owned by the caller.

        The first request downloads every column of the batch (one pass, one array per field);
        each array is handed out once without a further copy.  A column asked for again is
        downloaded again (the device copy is the source of truth — callers may mutate what
        they were given, e.g. apply_mask.py:184-185)."""
        if self.data is None:  # offloaded: the host arrays are the only copy
            return self._host[col].copy()
        if self._host is None:
            self._host = [np.empty((self.n_points,), dtype=self._np_dtype()) for _ in range(self.n_fields)]
            self._download(self._host)
        arr = self._host[col]
        if arr is None:
            arr = np.empty((self.n_points,), dtype=self._np_dtype())
            self._download([arr], first_col=col)
        else:
            self._host[col] = None
        return arr

    def host_column(self, col: int) -> np.ndarray:
        return self.take_column(col)


class HostPipeline:
    """Streaming regrid of host-resident fields (at_pipeline_*)."""

    def __init__(self, csr: CsrMatrix, chunk_fields: int = 64):
        require_cuda()
        self.csr = csr
        h = c_void_p()
        call("at_pipeline_create", csr.handle, int(chunk_fields), byref(h))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            call("at_pipeline_destroy", self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def regrid(self, fields_in: Sequence[np.ndarray], fields_out: Sequence[np.ndarray] | None = None) -> list[np.ndarray]:
        n = len(fields_in)
        n_src, n_tgt = self.csr.shape[1], self.csr.shape[0]
        ins = []
        for i, a in enumerate(fields_in):
            a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
            if a.size != n_src:
                raise ValueError(f"field {i} has {a.size} points, matrix expects {n_src}")
            ins.append(a)
        if fields_out is None:
            fields_out = [np.empty((n_tgt,), dtype=np.float32) for _ in range(n)]
        for i, a in enumerate(fields_out):
            if a.dtype != np.float32 or a.size != n_tgt or not a.flags.c_contiguous:
                raise ValueError(f"output field {i} must be contiguous float32[{n_tgt}]")
        in_ptrs = (c_void_p * n)(*[a.ctypes.data for a in ins])
        out_ptrs = (c_void_p * n)(*[a.ctypes.data for a in fields_out])
        call("at_pipeline_regrid", self._h, in_ptrs, out_ptrs, n)
        return list(fields_out)
