"""Spherical nearest-neighbour and mask functions — drop-in for `anemoi.transform.spatial`.

Same names, arguments, return types and assertions as the reference `spatial.py`; the
cKDTree build / query / query_ball_point calls and the per-point Python loop of
`cutout_mask` run on the GPU through libat_b200.so:

    reference spatial.py                      here
    --------------------                      ----
    cKDTree(points)                 96,396…   device.KnnIndex          (at_knn_create)
    .query(points, k, upper_bound)  96,396…   KnnIndex.query           (at_knn_query)
    .query_ball_point + set-union   533-534   KnnIndex.ball_mark + compact_mask
    cutout loop + Triangle3D        404-424   at_cutout_classify
    cropping_mask                   236-275   at_cropping_mask

`latlon_to_xyz` stays numpy on the host ON PURPOSE: numpy's float64 sin/cos are not
bit-identical to CUDA's, and nearest-neighbour indices are only bit-exact against cKDTree if
both searches see the same xyz.  It is O(n) preparation, not the search.

Tie policy: neighbours are ordered by (d², index).  On queries where two sources are at
exactly the same float64 d², cKDTree's pick depends on its traversal order; this
implementation returns the lowest index.  `nearest_grid_points(..., _return_ties=True)`
exposes the per-query tie flags.
"""

from __future__ import annotations

import logging
from typing import Any

import numpy as np
from numpy.typing import NDArray

from .constants import R_earth_km, radian
from .device import KnnIndex, call, compact_mask, cropping_mask_device, require_cuda, stream_ptr, to_device_f64, _ptr

LOG = logging.getLogger(__name__)

# np.dot on float64 3-vectors: OpenBLAS ddot accumulates with FMA on x86-64 hosts that have
# it; 0 switches at_cutout_classify to plain left-to-right products.
CUTOUT_DOT_MODE = 1


def xyz_to_latlon(x: NDArray[Any], y: NDArray[Any], z: NDArray[Any]) -> tuple[NDArray[Any], NDArray[Any]]:
    """Cartesian coordinates → latitude, longitude in degrees (spatial.py:109-129)."""
    return (
        np.rad2deg(np.arcsin(np.minimum(1.0, np.maximum(-1.0, z)))),
        np.rad2deg(np.arctan2(y, x)),
    )


def _latlon_to_xyz_serial(lat, lon, radius):
    phi = np.deg2rad(lat)
    lda = np.deg2rad(lon)
    cos_phi = np.cos(phi)
    x = cos_phi * np.cos(lda) * radius
    y = cos_phi * np.sin(lda) * radius
    z = np.sin(phi) * radius
    return x, y, z


_THREADED_TRIG_OK: bool | None = None
_TRIG_THREADS = None


def _threaded_trig_is_exact() -> bool:
    """numpy's ufuncs are elementwise, but their SIMD kernels treat array tails separately; check
    once per process that evaluating in slices gives the very same bits as one call, so the
    threaded path can never change a nearest-neighbour index."""
    global _THREADED_TRIG_OK
    if _THREADED_TRIG_OK is None:
        rng = np.random.default_rng(12345)
        lat, lon = rng.uniform(-90, 90, 20011), rng.uniform(-360, 720, 20011)
        whole = _latlon_to_xyz_serial(lat, lon, 1.0)
        cuts = [0, 1, 4, 9, 1000, 1003, 7777, 12345, 20011]
        parts = [_latlon_to_xyz_serial(lat[a:b], lon[a:b], 1.0) for a, b in zip(cuts[:-1], cuts[1:])]
        _THREADED_TRIG_OK = all(np.array_equal(np.concatenate([p[k] for p in parts]).view(np.uint64), whole[k].view(np.uint64)) for k in range(3))
    return _THREADED_TRIG_OK


def latlon_to_xyz(lat: NDArray[Any], lon: NDArray[Any], radius: float = 1.0) -> tuple[NDArray[Any], NDArray[Any], NDArray[Any]]:
    """Latitude, longitude in degrees → Cartesian coordinates on a sphere (spatial.py:132-167).

    numpy on the host, on purpose (module docstring).  Large inputs are evaluated in slices on a
    few host threads (numpy releases the GIL inside ufuncs) — the same numpy calls on the same
    elements, bit-identical to one call (verified once per process), several times faster for
    the millions of points of a global grid."""
    global _TRIG_THREADS
    lat_a, lon_a = np.asarray(lat), np.asarray(lon)
    n = lat_a.size
    if n < 500_000 or lat_a.ndim != 1 or lon_a.shape != lat_a.shape or lat_a.dtype != np.float64 or lon_a.dtype != np.float64 or not _threaded_trig_is_exact():
        return _latlon_to_xyz_serial(lat, lon, radius)
    import os
    from concurrent.futures import ThreadPoolExecutor

    # several ranks on one box share its cores
    local_ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    workers = max(1, min(8, (len(os.sched_getaffinity(0)) or 2) // local_ranks - 1))
    if workers == 1:
        return _latlon_to_xyz_serial(lat, lon, radius)
    if _TRIG_THREADS is None:
        _TRIG_THREADS = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="at-trig")
    x, y, z = (np.empty(n, dtype=np.float64) for _ in range(3))
    cuts = np.linspace(0, n, 4 * workers + 1).astype(np.int64)

    def work(ab):
        a, b = ab
        x[a:b], y[a:b], z[a:b] = _latlon_to_xyz_serial(lat_a[a:b], lon_a[a:b], radius)

    list(_TRIG_THREADS.map(work, zip(cuts[:-1], cuts[1:])))
    return x, y, z


class _Phases:
    """Wall-clock phases of a spatial function, logged when AT_B200_TIMING is set (each phase is
    closed with a device synchronisation, so it costs a little; off by default)."""

    def __init__(self, name: str):
        import os

        self.on = bool(os.environ.get("AT_B200_TIMING"))
        self.name, self.rows, self.t = name, [], None
        if self.on:
            self.mark(None)

    def mark(self, phase: str | None) -> None:
        if not self.on:
            return
        import time

        import torch

        torch.cuda.synchronize()
        now = time.perf_counter()
        if phase is not None:
            self.rows.append(f"{phase} {1e3 * (now - self.t):.1f}")
        self.t = now

    def report(self) -> None:
        if self.on:
            import os

            print(f"[{self.name} rank {os.environ.get('RANK', '0')}] " + " | ".join(self.rows) + " ms", flush=True)


def _check_latlon_arrays(lats, lons, global_lats, global_lons) -> None:
    assert global_lats.ndim == 1
    assert global_lons.ndim == 1
    assert lats.ndim == 1
    assert lons.ndim == 1
    assert global_lats.shape == global_lons.shape
    assert lats.shape == lons.shape


def _resolution(points_xyz, sharded: bool = False) -> float:
    """min over points of the distance to the 2nd nearest point (spatial.py:93-97).
    `sharded`: every rank scans its slice of the points, all-reduce(MIN)."""
    index = points_xyz if isinstance(points_xyz, KnnIndex) else KnnIndex(points_xyz)
    if not sharded:
        return index.min_nn_distance()
    from . import distributed as atd

    lo, hi = atd.shard_range(index.n, *atd.world())
    return atd.all_reduce_min(index.min_nn_distance(lo, hi - lo), device="cuda")


def _query_slice(n_q: int, sharded: bool) -> slice:
    """This rank's share of n_q independent queries: everything, or every world_size-th query
    (interleaved, so expensive far-away queries and cheap near ones mix on every rank)."""
    if not sharded:
        return slice(0, n_q)
    from . import distributed as atd

    rank, ws = atd.world()
    return slice(rank, n_q, ws)


def _xyz(lats, lons, sharded: bool):
    """xyz of the points: numpy arrays on the host — or, sharded, device tensors whose host
    evaluation was split over the ranks (`distributed.latlon_to_xyz_device`)."""
    if not sharded:
        return latlon_to_xyz(lats, lons)
    from . import distributed as atd

    return atd.latlon_to_xyz_device(lats, lons)


def _gather_queries(local, n_q: int, sharded: bool):
    if not sharded:
        return local
    from . import distributed as atd

    return atd.all_gather_strided(local, n_q)


def _distance_km_to_resolution(function: str, distance_km, lam_points, global_points) -> float:
    """spatial.py:100-106 — a number of km, or the resolution of "lam" / "global" / None."""
    if isinstance(distance_km, (int, float)):
        return distance_km / R_earth_km
    distance = _resolution({"lam": lam_points, "global": global_points, None: global_points}[distance_km])
    LOG.info(f"{function} using distance = {distance * R_earth_km} km")
    return distance


def cropping_mask(lats: NDArray[Any], lons: NDArray[Any], north: float, west: float, south: float, east: float) -> NDArray[Any]:
    """Points inside the box, bounds inclusive, longitudes tested at lon and lon ± 360."""
    require_cuda()
    mask = cropping_mask_device(np.asarray(lats, dtype=np.float64), np.asarray(lons, dtype=np.float64), north, west, south, east)
    return mask.cpu().numpy().astype(bool)


def _crop_box(lats, lons, distance):
    north, south = np.amax(lats), np.amin(lats)
    east, west = np.amax(lons), np.amin(lons)
    return (np.min([90.0, north + distance]), west - distance, np.max([-90.0, south - distance]), east + distance)


def cutout_mask(
    lats: NDArray[Any],
    lons: NDArray[Any],
    global_lats: NDArray[Any],
    global_lons: NDArray[Any],
    cropping_distance: float = 2.0,
    neighbours: int = 5,
    min_distance_km: int | float | None = None,
    max_distance_km: int | float | None = None,
    plot: str | None = None,
    _sharded: bool = False,
) -> NDArray[Any]:
    """Mask of the global points to KEEP around a LAM (True = outside the cutout).

    A global point is dropped when it is inside the LAM (a ray from the Earth's centre
    through it hits a triangle of its `neighbours` nearest LAM points), closer than
    `min_distance_km` to it, or farther than `max_distance_km` (spatial.py:294-440).
    """
    assert cropping_distance >= 0.0, "cropping_distance must be non-negative"
    assert min_distance_km is None or min_distance_km >= 0.0, "min_distance_km must be non-negative"
    assert max_distance_km is None or max_distance_km >= 0.0, "max_distance_km must be non-negative"
    assert neighbours > 0, "neighbours must be positive"
    _check_latlon_arrays(lats, lons, global_lats, global_lons)
    torch = require_cuda()

    effective_cropping_distance = cropping_distance
    if max_distance_km is not None:
        # make sure everything outside the crop box is farther than max_distance_km
        max_lat = max(abs(np.amax(lats)), abs(np.amin(lats)))
        R_earth_at_lat = R_earth_km * np.cos(np.deg2rad(max_lat))
        L_1_degree_arc_length_km = R_earth_at_lat * radian
        max_distance_degrees = max_distance_km / L_1_degree_arc_length_km
        effective_cropping_distance = max(cropping_distance, 1.1 * max_distance_degrees)

    mask = cropping_mask(global_lats, global_lons, *_crop_box(lats, lons, effective_cropping_distance))

    global_xyz = _xyz(global_lats[mask], global_lons[mask], _sharded)
    lam_xyz = _xyz(lats, lons, _sharded)
    n_lam, n_q = int(lam_xyz[0].shape[0]), int(global_xyz[0].shape[0])

    lam_index = KnnIndex(lam_xyz)
    if isinstance(min_distance_km, (int, float)):
        min_distance = min_distance_km / R_earth_km
    elif min_distance_km == "lam":
        min_distance = _resolution(lam_index, _sharded)
        LOG.info(f"cutout_mask using distance = {min_distance * R_earth_km} km")
    else:
        # None / "global" -> resolution of the (cropped) global points (spatial.py:388-393, 104)
        min_distance = _resolution(global_xyz, _sharded) if n_q > 0 else float("inf")
        LOG.info(f"cutout_mask using distance = {min_distance * R_earth_km} km")

    inside_lam = np.zeros((n_q,), dtype=bool)
    if n_q > 0:
        if neighbours > n_lam:
            # cKDTree pads with index n_lam and the reference then indexes lam_points with it
            raise IndexError(f"index {n_lam} is out of bounds for axis 0 with size {n_lam}")
        mine = _query_slice(n_q, _sharded)  # queries are independent: each rank classifies its share
        g = tuple(to_device_f64(a[mine]) for a in global_xyz)
        lam = tuple(to_device_f64(a) for a in lam_xyz)
        idx, dist, _ = lam_index.query(g, k=neighbours)
        n_mine = int(g[0].shape[0])
        out = torch.empty((n_mine,), dtype=torch.uint8, device=idx.device)
        max_distance = -1.0 if max_distance_km is None else max_distance_km / R_earth_km
        call(
            "at_cutout_classify",
            _ptr(lam[0]), _ptr(lam[1]), _ptr(lam[2]), n_lam,
            _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), n_mine,
            _ptr(idx), _ptr(dist), int(neighbours),
            float(min_distance), float(max_distance), int(CUTOUT_DOT_MODE),
            _ptr(out), stream_ptr(),
        )  # fmt: skip
        inside_lam = _gather_queries(out, n_q, _sharded).cpu().numpy().astype(bool)

    too_far_mask: bool | NDArray[Any] = False
    if isinstance(max_distance_km, (int, float)):
        too_far_mask = ~mask.copy()  # everything outside the cropping area is too far

    mask[mask] = inside_lam
    mask[too_far_mask] = True
    mask = ~mask

    if plot:
        raise NotImplementedError("plotting is not part of this package")
    return mask


def thinning_mask(
    lats: NDArray[Any],
    lons: NDArray[Any],
    global_lats: NDArray[Any],
    global_lons: NDArray[Any],
    cropping_distance: float = 2.0,
    _sharded: bool = False,
) -> NDArray[Any]:
    """Indices of the LAM points closest to each (cropped) global point (spatial.py:443-503)."""
    _check_latlon_arrays(lats, lons, global_lats, global_lons)
    require_cuda()
    ph = _Phases("thinning_mask")
    mask = cropping_mask(global_lats, global_lons, *_crop_box(lats, lons, cropping_distance))
    ph.mark("crop")
    global_xyz = _xyz(global_lats[mask], global_lons[mask], _sharded)
    ph.mark("xyz global")
    lam_xyz = _xyz(lats, lons, _sharded)
    ph.mark("xyz lam")
    index = KnnIndex(lam_xyz)
    ph.mark("build")
    n_q = int(global_xyz[0].shape[0])
    mine = _query_slice(n_q, _sharded)
    idx, _, _ = index.query(tuple(a[mine] for a in global_xyz), k=1)
    ph.mark("query")
    out = _gather_queries(idx[:, 0].contiguous(), n_q, _sharded).cpu().numpy()
    ph.mark("gather + D2H")
    ph.report()
    return out


def global_on_lam_mask(
    lats: NDArray[Any],
    lons: NDArray[Any],
    global_lats: NDArray[Any],
    global_lons: NDArray[Any],
    distance_km: float | None = None,
) -> NDArray[Any]:
    """Sorted indices of the global points within `distance` of any LAM point (spatial.py:506-536)."""
    _check_latlon_arrays(lats, lons, global_lats, global_lons)
    require_cuda()
    global_index = KnnIndex(latlon_to_xyz(global_lats, global_lons))
    lam_xyz = latlon_to_xyz(lats, lons)
    if isinstance(distance_km, (int, float)):
        distance = distance_km / R_earth_km
    else:
        source = KnnIndex(lam_xyz) if distance_km == "lam" else global_index
        distance = _resolution(source)
        LOG.info(f"global_on_lam_mask using distance = {distance * R_earth_km} km")
    mark = global_index.ball_mark(lam_xyz, distance)
    indices = compact_mask(mark).cpu().numpy()
    if indices.size == 0:
        return np.array(sorted(set()))  # the reference's empty result is a float64 array
    return indices


def outline(lats: NDArray[Any], lons: NDArray[Any], neighbours: int = 5) -> list[int]:
    """Indices of the outline points of a grid (spatial.py:539-584): the points that are not
    inside any triangle of the fan spanned by their own nearest neighbours.

    The fan is built from the neighbours in cKDTree's order; where several neighbours are at
    exactly the same float64 distance (regular grids) that order is cKDTree's traversal order
    and this implementation's is (distance, index) — the module's tie policy — so on such grids
    the outline can differ from the reference's in the tied points."""
    torch = require_cuda()
    xyz = latlon_to_xyz(lats, lons)
    n = int(xyz[0].shape[0])
    if n == 0:
        return []
    if neighbours > n:
        raise IndexError(f"index {n} is out of bounds for axis 0 with size {n}")
    index = KnnIndex(xyz)
    pts = tuple(to_device_f64(a) for a in xyz)
    idx, dist, _ = index.query(pts, k=neighbours)
    inside = torch.empty((n,), dtype=torch.uint8, device=idx.device)
    call("at_outline_classify", _ptr(pts[0]), _ptr(pts[1]), _ptr(pts[2]), n, _ptr(idx), _ptr(dist), int(neighbours), int(CUTOUT_DOT_MODE), _ptr(inside), stream_ptr())
    return np.nonzero(inside.cpu().numpy() == 0)[0].tolist()


def nearest_grid_points(
    source_latitudes: NDArray[Any],
    source_longitudes: NDArray[Any],
    target_latitudes: NDArray[Any],
    target_longitudes: NDArray[Any],
    max_distance: float | None = None,
    num_neighbours_to_return: int = 1,
    return_distances: bool = False,
    _as_device: bool = False,
    _return_ties: bool = False,
) -> NDArray[Any] | tuple[NDArray[Any], NDArray[Any]]:
    """Nearest source grid points of each target point (spatial.py:587-635).

    Indices have shape [n_target] for one neighbour, [n_target, k] otherwise, ascending by
    distance; with `max_distance` a target with no source strictly inside the bound gets
    index n_source and distance inf, as cKDTree reports misses.
    """
    require_cuda()
    source_xyz = latlon_to_xyz(source_latitudes, source_longitudes)
    target_xyz = latlon_to_xyz(target_latitudes, target_longitudes)
    index = KnnIndex(source_xyz)
    k = int(num_neighbours_to_return)
    ub = float("inf") if max_distance is None else float(max_distance)
    idx, dist, ties = index.query(target_xyz, k=k, distance_upper_bound=ub, want_ties=_return_ties)
    if k == 1:
        idx, dist = idx[:, 0], dist[:, 0]
    if _as_device:
        return idx
    indices, distances = idx.cpu().numpy(), dist.cpu().numpy()
    if _return_ties:
        return indices, distances, ties.cpu().numpy()
    if return_distances:
        return indices, distances
    return indices
