"""Oracle: restatement of the reference's spatial functions on scipy.spatial.cKDTree.

TEST INFRASTRUCTURE — see oracle/__init__.py.

Each function follows the reference function of the same name in
src/anemoi/transform/spatial.py (line ranges in the docstrings).  The k-d tree itself is
scipy's (a transitive dependency of the reference, installed here and on the GPU box), so
neighbour indices, distances and ball queries are the reference's own arithmetic.  Two
extra functions document what that arithmetic is:

    knn_bruteforce   d² = ((dx·dx)+(dy·dy)+(dz·dz)) in float64, neighbours ordered by
                     (d², index) — bitwise cKDTree's distances; shows which queries have
                     exact ties (where cKDTree's pick depends on its traversal order)
    cutout_mask_vectorised   the cutout loop with unfused dot products, whole-array

PINNED: against the UNMODIFIED reference imported in the build container
(oracle/make_golden.py → tests/golden/spatial_*.npz) and against the reference's own
known-answer tests tests/test_spatial.py:18-145 (tests/test_oracle_spatial.py).
"""

from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

R_earth_km = 6371229.0 / 1000
radian = np.pi / 180.0


def latlon_to_xyz(lat, lon, radius: float = 1.0):
    """spatial.py:132-167."""
    phi = np.deg2rad(lat)
    lda = np.deg2rad(lon)
    cos_phi = np.cos(phi)
    cos_lda = np.cos(lda)
    sin_phi = np.sin(phi)
    sin_lda = np.sin(lda)
    return cos_phi * cos_lda * radius, cos_phi * sin_lda * radius, sin_phi * radius


def xyz_to_latlon(x, y, z):
    """spatial.py:109-129."""
    return np.rad2deg(np.arcsin(np.minimum(1.0, np.maximum(-1.0, z)))), np.rad2deg(np.arctan2(y, x))


def _points(lat, lon) -> np.ndarray:
    return np.array(latlon_to_xyz(lat, lon)).transpose()


def resolution(points: np.ndarray) -> float:
    """spatial.py:93-97: smallest distance of a point to its 2nd nearest point (itself = 1st)."""
    distances, _ = cKDTree(points).query(points, k=2)
    return np.min(distances[:, 1])


def distance_km_to_resolution(distance_km, lam_points, global_points) -> float:
    """spatial.py:100-106."""
    if isinstance(distance_km, (int, float)):
        return distance_km / R_earth_km
    return resolution({"lam": lam_points, "global": global_points, None: global_points}[distance_km])


def cropping_mask(lats, lons, north, west, south, east):
    """spatial.py:236-275."""
    return (
        (lats >= south)
        & (lats <= north)
        & (((lons >= west) & (lons <= east)) | ((lons >= west + 360) & (lons <= east + 360)) | ((lons >= west - 360) & (lons <= east - 360)))
    )


def _crop(lats, lons, global_lats, global_lons, distance):
    north, south, east, west = np.amax(lats), np.amin(lats), np.amax(lons), np.amin(lons)
    return cropping_mask(global_lats, global_lons, np.min([90.0, north + distance]), west - distance, np.max([-90.0, south - distance]), east + distance)


def nearest_grid_points(source_latitudes, source_longitudes, target_latitudes, target_longitudes, max_distance=None, num_neighbours_to_return=1, return_distances=False):
    """spatial.py:587-635."""
    source_points = _points(source_latitudes, source_longitudes)
    target_points = _points(target_latitudes, target_longitudes)
    if max_distance is None:
        distances, indices = cKDTree(source_points).query(target_points, k=num_neighbours_to_return)
    else:
        distances, indices = cKDTree(source_points).query(target_points, k=num_neighbours_to_return, distance_upper_bound=max_distance)
    if return_distances:
        return indices, distances
    return indices


def thinning_mask(lats, lons, global_lats, global_lons, cropping_distance: float = 2.0):
    """spatial.py:443-503."""
    mask = _crop(lats, lons, global_lats, global_lons, cropping_distance)
    global_points = _points(global_lats[mask], global_lons[mask])
    _, indices = cKDTree(_points(lats, lons)).query(global_points, k=1)
    return indices


def global_on_lam_mask(lats, lons, global_lats, global_lons, distance_km=None):
    """spatial.py:506-536 (set-union written as unique-of-concatenation: same sorted result)."""
    global_points = _points(global_lats, global_lons)
    lam_points = _points(lats, lons)
    distance = distance_km_to_resolution(distance_km, lam_points, global_points)
    lists = cKDTree(global_points).query_ball_point(lam_points, distance)
    flat = [i for sub in lists for i in sub]
    return np.array(sorted(set(flat)))


# ---- cutout -----------------------------------------------------------------------------
def triangle_intersect(v0, v1, v2, ray_origin, ray_direction) -> bool:
    """Triangle3D.intersect, spatial.py:189-233 (Möller–Trumbore, np.cross / np.dot)."""
    epsilon = 0.0000001
    h = np.cross(ray_direction, v2 - v0)
    a = np.dot(v1 - v0, h)
    if -epsilon < a < epsilon:
        return False
    f = 1.0 / a
    s = ray_origin - v0
    u = f * np.dot(s, h)
    if u < 0.0 or u > 1.0:
        return False
    q = np.cross(s, v1 - v0)
    v = f * np.dot(ray_direction, q)
    if v < 0.0 or u + v > 1.0:
        return False
    t = f * np.dot(v2 - v0, q)
    return bool(t > epsilon)


def _cutout_setup(lats, lons, global_lats, global_lons, cropping_distance, min_distance_km, max_distance_km):
    assert cropping_distance >= 0.0
    assert global_lats.ndim == 1 and global_lons.ndim == 1 and lats.ndim == 1 and lons.ndim == 1
    assert global_lats.shape == global_lons.shape and lats.shape == lons.shape
    effective = cropping_distance
    if max_distance_km is not None:
        max_lat = max(abs(np.amax(lats)), abs(np.amin(lats)))
        L = R_earth_km * np.cos(np.deg2rad(max_lat)) * radian
        effective = max(cropping_distance, 1.1 * (max_distance_km / L))
    mask = _crop(lats, lons, global_lats, global_lons, effective)
    global_points = _points(global_lats[mask], global_lons[mask])
    lam_points = _points(lats, lons)
    min_distance = distance_km_to_resolution(min_distance_km, lam_points, global_points)
    return mask, global_points, lam_points, min_distance


def _cutout_finish(mask, inside_lam, max_distance_km):
    too_far_mask = False
    if isinstance(max_distance_km, (int, float)):
        too_far_mask = ~mask.copy()
    mask[mask] = inside_lam
    mask[too_far_mask] = True
    return ~mask


def cutout_mask(lats, lons, global_lats, global_lons, cropping_distance=2.0, neighbours=5, min_distance_km=None, max_distance_km=None, distances=None, indices=None):
    """spatial.py:294-440 with the per-point Python loop kept literal (small inputs only).

    `distances` / `indices` may be supplied to classify with a given neighbour order
    (used to check the classification independently of tie order)."""
    mask, global_points, lam_points, min_distance = _cutout_setup(lats, lons, global_lats, global_lons, cropping_distance, min_distance_km, max_distance_km)
    if distances is None:
        distances, indices = cKDTree(lam_points).query(global_points, k=neighbours)
    if neighbours == 1:
        distances, indices = distances.reshape(-1, 1), indices.reshape(-1, 1)
    zero = np.array([0.0, 0.0, 0.0])
    inside_lam = []
    for global_point, distance, index in zip(global_points, distances, indices):
        inside = False
        for j in range(neighbours):
            inside = triangle_intersect(lam_points[index[j]], lam_points[index[(j + 1) % neighbours]], lam_points[index[(j + 2) % neighbours]], zero, global_point)
            if inside:
                break
        close = np.min(distance) <= min_distance
        too_far = False
        if max_distance_km is not None:
            too_far = np.min(distance) > (max_distance_km / R_earth_km)
        inside_lam.append(bool(inside or close or too_far))
    return _cutout_finish(mask, np.array(inside_lam, dtype=bool), max_distance_km)


def outline(lats, lons, neighbours=5, indices=None):
    """spatial.py:539-584 with the per-point Python loop kept literal (small inputs only).
    `indices` may be supplied to build the fans from a given neighbour order (ties)."""
    grid_points = _points(lats, lons)
    if indices is None:
        _, indices = cKDTree(grid_points).query(grid_points, k=neighbours)
    zero = np.array([0.0, 0.0, 0.0])
    outside = []
    for i, (point, index) in enumerate(zip(grid_points, indices)):
        inside = False
        for j in range(1, neighbours):
            inside = triangle_intersect(grid_points[index[j]], grid_points[index[(j + 1) % neighbours]], grid_points[index[(j + 2) % neighbours]], zero, point)
            if inside:
                break
        if not inside:
            outside.append(i)
    return outside


def _vcross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


def _vdot(a, b):
    return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]


def cutout_mask_vectorised(lats, lons, global_lats, global_lons, cropping_distance=2.0, neighbours=5, min_distance_km=None, max_distance_km=None):
    """The same classification, whole-array, with unfused dot products (for mid-size grids)."""
    mask, global_points, lam_points, min_distance = _cutout_setup(lats, lons, global_lats, global_lons, cropping_distance, min_distance_km, max_distance_km)
    distances, indices = cKDTree(lam_points).query(global_points, k=neighbours)
    if neighbours == 1:
        distances, indices = distances.reshape(-1, 1), indices.reshape(-1, 1)
    eps = 0.0000001
    n = global_points.shape[0]
    inside = np.zeros(n, dtype=bool)
    d = global_points
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for j in range(neighbours):
            v0 = lam_points[indices[:, j]]
            v1 = lam_points[indices[:, (j + 1) % neighbours]]
            v2 = lam_points[indices[:, (j + 2) % neighbours]]
            e2, e1 = v2 - v0, v1 - v0
            h = _vcross(d, e2)
            a = _vdot(e1, h)
            ok = ~((-eps < a) & (a < eps))
            f = 1.0 / a
            s = 0.0 - v0
            u = f * _vdot(s, h)
            ok &= ~((u < 0.0) | (u > 1.0))
            q = _vcross(s, e1)
            v = f * _vdot(d, q)
            ok &= ~((v < 0.0) | (u + v > 1.0))
            t = f * _vdot(e2, q)
            ok &= t > eps
            inside |= ok
    dmin = distances.min(axis=1)
    close = dmin <= min_distance
    too_far = np.zeros(n, dtype=bool) if max_distance_km is None else dmin > (max_distance_km / R_earth_km)
    return _cutout_finish(mask, inside | close | too_far, max_distance_km)


# ---- what cKDTree computes, spelled out ---------------------------------------------------
def knn_bruteforce(source_points: np.ndarray, target_points: np.ndarray, k: int = 1, chunk: int = 256):
    """Exact k-NN by exhaustive search: d² = ((dx·dx)+(dy·dy))+(dz·dz) in float64, unfused,
    neighbours ordered by (d², index).  Returns (indices [n,k], distances [n,k],
    tie [n] bool: the k-th and (k+1)-th d² are equal or two selected d² are equal)."""
    n = target_points.shape[0]
    idx = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=np.float64)
    tie = np.zeros(n, dtype=bool)
    sx, sy, sz = source_points[:, 0], source_points[:, 1], source_points[:, 2]
    for i0 in range(0, n, chunk):
        t = target_points[i0 : i0 + chunk]
        dx = sx[None, :] - t[:, 0:1]
        dy = sy[None, :] - t[:, 1:2]
        dz = sz[None, :] - t[:, 2:3]
        d2 = (dx * dx + dy * dy) + dz * dz
        order = np.argsort(d2, axis=1, kind="stable")  # stable: ties keep the lower index first
        kk = min(k + 1, d2.shape[1])
        top = order[:, :kk]
        dtop = np.take_along_axis(d2, top, axis=1)
        idx[i0 : i0 + chunk] = top[:, :k]
        dist[i0 : i0 + chunk] = np.sqrt(dtop[:, :k])
        tie[i0 : i0 + chunk] = (np.diff(dtop, axis=1) == 0).any(axis=1)
    return idx, dist, tie
