"""The field I/O engine (csrc/hostio.cu): pinned pool, upload / download, streamed regrid —
and the drop-in filters on top of it (FieldList of ordinary numpy fields in, numpy out)."""

import gc

import numpy as np
import pytest
from conftest import assert_same_values
from scipy.sparse import csr_array

from anemoi_transform_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _matrix(dtype=np.float32):
    t_lat, t_lon = syn.octahedral(48)
    d, i, p, shape = syn.bilinear_matrix(2.0, t_lat, t_lon)
    return d.astype(dtype), i, p, shape, (t_lat, t_lon)


def test_pinned_pool_recycles_blocks_and_rejects_double_free(cuda):
    from anemoi_transform_b200 import _cabi
    from anemoi_transform_b200.device import pinned_empty, pinned_pool_stats

    a = pinned_empty((1000,), np.float32)
    a[:] = 7.0
    addr = a.__array_interface__["data"][0]
    in_use, reserved = pinned_pool_stats()
    assert in_use >= 4000 and reserved >= in_use
    view = a[10:20]  # views keep the block alive
    del a
    gc.collect()
    assert pinned_pool_stats()[0] == in_use and view[0] == 7.0
    del view
    gc.collect()
    assert pinned_pool_stats()[0] == in_use - 4096
    b = pinned_empty((1000,), np.float32)  # same size class: the block comes back
    assert b.__array_interface__["data"][0] == addr
    lib = _cabi.load()
    import ctypes

    p = ctypes.c_void_p()
    assert lib.at_pinned_alloc(64, ctypes.byref(p)) == 0
    assert lib.at_pinned_free(p) == 0
    assert lib.at_pinned_free(p) != 0 and b"double free" in lib.at_last_error()
    assert lib.at_pinned_free(ctypes.c_void_p(12345)) != 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n_fields,n_points", [(1, 5), (3, 1000), (37, 65160), (130, 40320)])
def test_upload_download_round_trip(cuda, dtype, n_fields, n_points):
    """Pageable and page-locked arrays, in and out, through at_hostio_upload / at_hostio_download."""
    from anemoi_transform_b200.device import DeviceBatch, pinned_empty

    rng = np.random.default_rng(n_fields)
    fields = [rng.standard_normal(n_points).astype(dtype) for _ in range(n_fields)]
    fields[0][::7] = np.nan
    batch = DeviceBatch.from_host_fields(fields)
    assert batch.data.shape == (n_points, (n_fields + 3) // 4 * 4)
    assert_same_values(batch.data[:, :n_fields].cpu().numpy().T, np.stack(fields), "upload")
    assert_same_values(batch.to_host_fields(), np.stack(fields), "pageable download")
    # page-locked inputs are DMA-ed in place
    pinned = []
    for f in fields:
        a = pinned_empty((n_points,), dtype)
        a[:] = f
        pinned.append(a)
    again = DeviceBatch.from_host_fields(pinned)
    assert cuda.equal(again.data[:, :n_fields].nan_to_num(7.0), batch.data[:, :n_fields].nan_to_num(7.0))
    # take_column: pool arrays, handed out once; a second request downloads again
    for j in (0, n_fields - 1, n_fields // 2):
        first = batch.take_column(j)
        assert_same_values(first, fields[j], f"column {j}")
        first[:] = -1.0  # the caller owns it
        assert_same_values(batch.take_column(j), fields[j], f"column {j} again")


def test_take_column_downloads_a_window_not_the_batch(cuda):
    """reading one field of a large batch must not download every column."""
    from anemoi_transform_b200.device import DeviceBatch, pinned_pool_stats

    n_points, n_fields = 1 << 16, 256
    data = cuda.arange(n_points * n_fields, dtype=cuda.float32, device="cuda").reshape(n_points, n_fields)
    batch = DeviceBatch(data, n_fields)
    batch.window_bytes = 8 * n_points * 4  # 8 fields per window
    gc.collect()
    before = pinned_pool_stats()[0]
    col = batch.take_column(100)
    assert np.array_equal(col, data[:, 100].cpu().numpy())
    held = pinned_pool_stats()[0] - before
    assert held <= 8 * (n_points * 4 + 4096), held
    assert batch._host[100] is None and batch._host[101] is not None and batch._host[99] is None and batch._host[108] is None


@pytest.mark.parametrize("mdtype,xdtype", [(np.float32, np.float32), (np.float64, np.float32), (np.float32, np.float64), (np.float64, np.float64)])
@pytest.mark.parametrize("n_fields", [1, 6, 64, 131])
def test_streamed_regrid_matches_scipy(cuda, mdtype, xdtype, n_fields):
    from anemoi_transform_b200 import _cabi
    from anemoi_transform_b200.device import CsrMatrix, StreamedRegrid

    d, i, p, shape, _ = _matrix(mdtype)
    m = csr_array((d, i, p), shape=shape)
    csr = CsrMatrix(d, i, p, shape)
    fields = [syn.synthetic_field("t", shape[1], s, 0.001 if s % 3 == 0 else 0.0).astype(xdtype) for s in range(n_fields)]
    want = np.stack([m @ f for f in fields])
    ydt = csr.result_dtype(cuda.float32 if xdtype == np.float32 else cuda.float64)
    for keep in (True, False):
        batch = StreamedRegrid(_cabi.HOSTIO_SPMM, csr, None, shape[0], ydt, fields, keep_resident=keep, to_host=True).join()
        got = np.stack([batch.take_column(j) for j in range(n_fields)])
        assert got.dtype == want.dtype
        assert_same_values(got, want, f"streamed keep={keep}")
        if keep:
            assert_same_values(batch.data[:, :n_fields].cpu().numpy().T, want, "resident result")
    only_resident = StreamedRegrid(_cabi.HOSTIO_SPMM, csr, None, shape[0], ydt, fields, keep_resident=True, to_host=False).join()
    assert_same_values(only_resident.to_host_fields(), want, "resident only")


def test_streamed_regrid_rejects_a_wrong_grid(cuda):
    from anemoi_transform_b200 import _cabi
    from anemoi_transform_b200.device import CsrMatrix, StreamedRegrid

    d, i, p, shape, _ = _matrix()
    csr = CsrMatrix(d, i, p, shape)
    with pytest.raises(ValueError, match="dimension mismatch"):
        StreamedRegrid(_cabi.HOSTIO_SPMM, csr, None, shape[0], cuda.float32, [np.zeros(shape[1] - 1, np.float32)], True, True).join()


def _fieldlist(values, lat, lon, params=None):
    from anemoi_transform_b200 import ekd

    return ekd.from_source(
        "list-of-dicts",
        [dict(param=(params[k] if params else "t"), levelist=850, step=k, values=v, latitudes=lat, longitudes=lon) for k, v in enumerate(values)],
    )


def test_regrid_filter_streams_numpy_fields_for_all_three_interpolators(cuda, tmp_path):
    """FieldList of ordinary numpy fields -> RegridFilter.forward -> to_numpy of every output:
    matrix, nearest and mask variants equal scipy / numpy indexing bit for bit."""
    from scipy.spatial import cKDTree

    from anemoi_transform_b200.filters import create_filter_by_name as F
    from oracle import spatial as osp

    d, i, p, shape, (t_lat, t_lon) = _matrix()
    s_lat, s_lon = syn.regular_latlon(2.0)
    path = str(tmp_path / "m.npz")
    syn.save_regrid_npz(path, d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    m = csr_array((d, i, p), shape=shape)
    values = [syn.synthetic_field("t", shape[1], s, 0.002 if s == 1 else 0.0) for s in range(21)]
    values[5] = values[5].astype(np.float64)  # each field keeps its own dtype
    data = _fieldlist(values, s_lat, s_lon)

    out = F("regrid", matrix=path).forward(data)
    assert len(out) == len(values)
    for k, f in enumerate(out):
        got = f.to_numpy()
        assert got.dtype == (m @ values[k]).dtype
        assert_same_values(got, m @ values[k], f"matrix field {k}")
        assert np.array_equal(f.grid_points()[0], t_lat)

    out = F("regrid", method="nearest", in_grid=dict(latitudes=s_lat, longitudes=s_lon), out_grid=dict(latitudes=t_lat, longitudes=t_lon)).forward(data)
    s_xyz, q_xyz = osp.latlon_to_xyz(s_lat, s_lon), osp.latlon_to_xyz(t_lat, t_lon)
    dist, idx = cKDTree(np.array(s_xyz).T).query(np.array(q_xyz).T, k=1)
    for k, f in enumerate(out):
        got = f.to_numpy()
        # ties may pick another source at the same distance: compare through the distances
        same = got == values[k][idx]
        both_nan = np.isnan(got) & np.isnan(values[k][idx])
        assert (same | both_nan).mean() > 0.99

    mask = np.sort(np.random.default_rng(3).choice(shape[1], 5000, replace=False))
    mpath = str(tmp_path / "mask.npz")
    np.savez(mpath, mask=mask)
    out = F("regrid", mask=mpath).forward(data)
    for k, f in enumerate(out):
        assert_same_values(f.to_numpy(), values[k][mask], f"mask field {k}")
        assert np.array_equal(f.grid_points()[1], s_lon[mask])
    bad = str(tmp_path / "bad.npz")
    np.savez(bad, mask=np.array([0, shape[1]]))
    with pytest.raises(IndexError):
        F("regrid", mask=bad).forward(data)


def test_pipeline_keeps_intermediates_on_the_device_and_prefetches_the_last(cuda, tmp_path):
    from anemoi_transform_b200.fields import device_column_of
    from anemoi_transform_b200.filters import create_filter_by_name as F
    from anemoi_transform_b200.source import FieldListSource

    d, i, p, shape, (t_lat, t_lon) = _matrix()
    s_lat, s_lon = syn.regular_latlon(2.0)
    path = str(tmp_path / "m.npz")
    syn.save_regrid_npz(path, d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    m = csr_array((d, i, p), shape=shape)
    values = [syn.synthetic_field("t", shape[1], s) for s in range(8)]
    data = _fieldlist(values, s_lat, s_lon)
    regrid, rescale = F("regrid", matrix=path), F("rescale", param="t", scale=2.0, offset=1.0)
    # stand-alone: results are on their way to the host when forward returns
    out = regrid.forward(data)
    batch, _ = device_column_of(out[0])
    assert batch._host is not None and all(a is not None for a in batch._host)
    # inside a pipeline the regrid leaves them in HBM; the last filter prefetches
    from anemoi_transform_b200.filter import Filter

    seen = {}

    class Spy(Filter):
        def forward(self, fl):
            b, _ = device_column_of(fl[0])
            seen["intermediate_host"] = b._host
            return fl

    result = (FieldListSource(dataset=data) | regrid | Spy() | rescale).forward(None)
    assert seen["intermediate_host"] is None
    for k, f in enumerate(result):
        want = (m @ values[k]) * np.float32(2.0) + np.float32(1.0)
        assert_same_values(f.to_numpy(), want, f"pipeline field {k}")


def test_fused_pipeline_on_a_wrong_grid_raises_like_scipy(cuda, tmp_path):
    """`regrid | clip` fed fields of the wrong grid must raise, not read out of bounds."""
    from anemoi_transform_b200.device import CsrMatrix, Epilogue
    from anemoi_transform_b200 import _cabi
    from anemoi_transform_b200.filters import create_filter_by_name as F
    from anemoi_transform_b200.source import FieldListSource

    d, i, p, shape, (t_lat, t_lon) = _matrix()
    s_lat, s_lon = syn.regular_latlon(2.0)
    path = str(tmp_path / "m.npz")
    syn.save_regrid_npz(path, d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    w_lat, w_lon = syn.regular_latlon(4.0)
    wrong = _fieldlist([np.ones(w_lat.size, np.float32) for _ in range(4)], w_lat, w_lon)
    pipe = FieldListSource(dataset=wrong) | F("regrid", matrix=path) | F("clip", param="t", minimum=0.0)
    with pytest.raises(ValueError, match="dimension mismatch"):
        [f.to_numpy() for f in pipe.forward(None)]
    csr = CsrMatrix(d, i, p, shape)
    epi = Epilogue([(_cabi.EPI_PLAIN, 0, 4, 0)], [(0, 0, 0, 0)] * 4)
    with pytest.raises(ValueError, match="dimension mismatch"):
        epi.apply_fused(csr, cuda.zeros((shape[1] + 5, 4), device="cuda"))
    with pytest.raises(ValueError, match="input columns"):
        Epilogue([(_cabi.EPI_PLAIN, 0, 8, 0)], [(0, 0, 0, 0)] * 8).apply_fused(csr, cuda.zeros((shape[1], 4), device="cuda"))


def test_empty_query_sets_return_empty_results(cuda):
    """zero-length CUDA tensors have null data pointers; an empty query set is not an error."""
    from anemoi_transform_b200.device import KnnIndex

    s_xyz = tuple(np.random.default_rng(0).standard_normal((3, 1000)))
    index = KnnIndex(s_xyz)
    empty = tuple(np.empty((0,), dtype=np.float64) for _ in range(3))
    idx, dist, ties = index.query(empty, k=3, want_ties=True)
    assert idx.shape == (0, 3) and dist.shape == (0, 3) and ties.shape == (0,)
    mark = index.ball_mark(empty, 0.1)
    assert int(mark.sum()) == 0


def test_mixed_dtype_fieldlist_keeps_each_fields_dtype(cuda):
    """numpy processes each field in its own dtype; a float32 field next to a float64
    one must come back float32 with float32 arithmetic."""
    from anemoi_transform_b200.filters import create_filter_by_name as F

    lat, lon = syn.regular_latlon(4.0)
    rng = np.random.default_rng(5)
    values = [rng.standard_normal(lat.size).astype(np.float32) * 10, rng.standard_normal(lat.size) * 10, rng.standard_normal(lat.size).astype(np.float32)]
    data = _fieldlist(values, lat, lon)
    out = F("clip", param="t", minimum=-1.5, maximum=2.25).forward(data)
    for v, f in zip(values, out):
        got = f.to_numpy()
        assert got.dtype == v.dtype
        assert_same_values(got.reshape(-1), np.clip(v, -1.5, 2.25), "clip")
    out = F("rescale", param="t", scale=1.1, offset=0.3).forward(data)
    for v, f in zip(values, out):
        got = f.to_numpy()
        assert got.dtype == v.dtype
        assert_same_values(got.reshape(-1), v * v.dtype.type(1.1) + v.dtype.type(0.3), "rescale")


def test_config3_grids_through_the_plugin_call_at_scale(cuda, tmp_path):
    """Config 3's grids (0.25° → N320-shaped) through the call bench.py times as `e2e`: 600
    ordinary numpy fields (ragged last chunk, one float64 field, NaNs) → RegridFilter.forward →
    to_numpy() of every output; every field of a sample bitwise scipy's, every other checked
    through a checksum against the resident result."""
    from anemoi_transform_b200.fields import device_column_of
    from anemoi_transform_b200.filters import create_filter_by_name as F

    t_lat, t_lon = syn.n320_like()
    s_lat, s_lon = syn.regular_latlon(0.25)
    d, i, p, shape = syn.bilinear_matrix(0.25, t_lat, t_lon)
    path = str(tmp_path / "c3.npz")
    syn.save_regrid_npz(path, d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    m = csr_array((d, i, p), shape=shape)
    n = 603
    rng = np.random.default_rng(11)
    base = rng.standard_normal((16, shape[1]), dtype=np.float32)
    values = [np.array(base[k % 16]) * np.float32(1 + k) for k in range(n)]
    values[7][::1001] = np.nan
    values[300] = values[300].astype(np.float64)
    data = _fieldlist(values, s_lat, s_lon)
    out = F("regrid", matrix=path).forward(data)
    arrays = [f.to_numpy() for f in out]
    assert len(arrays) == n and all(a.shape == (shape[0],) for a in arrays)
    for k in (0, 7, 59, 60, 61, 300, 599, 602):
        assert arrays[k].dtype == (m @ values[k]).dtype
        assert_same_values(arrays[k], m @ values[k], f"field {k}")
    # linearity: field k is (1 + k) x base[k % 16] exactly in float32 wherever no rounding differs;
    # the host copy of every field equals the resident column it was downloaded from
    batch, col0 = device_column_of(out[0])
    resident = batch.data[:, : batch.n_fields].cpu().numpy()
    cols = [device_column_of(f)[1] for f in out if device_column_of(f)[0] is batch]
    for j, c in enumerate(cols[:: 37]):
        k = [kk for kk, f in enumerate(out) if device_column_of(f)[0] is batch][:: 37][j]
        assert_same_values(arrays[k], resident[:, c], f"host copy of field {k}")
