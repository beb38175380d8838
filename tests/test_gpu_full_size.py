"""BASELINE configs 4 and 5 at their full sizes, checked through properties that do not need the
oracle to process the whole problem (config 2 and 3 at full size: test_gpu_knn.py /
test_gpu_spmm.py).  What is exact here stays exact: bitwise fused = unfused, scipy on sampled
columns, cKDTree on sampled points, the vectorised oracle on a sampled subset of global points.
"""

import numpy as np
import pytest
from conftest import assert_same_values
from scipy.sparse import csr_array
from scipy.spatial import cKDTree

from anemoi_transform_b200 import synthetic as syn
from oracle import spatial as osp

pytestmark = pytest.mark.gpu


def test_config4_full_size_regrid_and_fused_epilogue(cuda):
    """O1280 (6,599,680) → N320-shaped, 12 nonzeros per row (inverse-distance weights of the 12
    nearest sources, found by the device kNN), 256 fields."""
    from anemoi_transform_b200 import _cabi, spatial
    from anemoi_transform_b200.device import CsrMatrix, Epilogue, KnnIndex

    s, t = syn.octahedral(1280), syn.n320_like()
    sx, tx = spatial.latlon_to_xyz(*s), spatial.latlon_to_xyz(*t)
    knn = KnnIndex(sx)
    idx, dist, _ = knn.query(tuple(cuda.from_numpy(a).cuda() for a in tx), k=12)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    # the 12 neighbours of a sample of targets are the brute-force ones, in ascending order
    pick = np.random.default_rng(0).choice(idx.shape[0], 200, replace=False)
    S = np.array(sx).T
    for j in pick:
        d = np.sqrt(((S - np.array([tx[0][j], tx[1][j], tx[2][j]])) ** 2).sum(axis=1))
        want = np.sort(d)[:12]
        assert np.array_equal(np.sort(dist[j]), dist[j]) and np.allclose(dist[j], want, rtol=0, atol=1e-15)
    d, i, p, shape = syn.knn_matrix(idx, dist, sx[0].size)
    assert shape == (542_080, 6_599_680) and d.size == 12 * 542_080
    del knn
    csr = CsrMatrix(d, i, p, shape)
    assert csr.uniform_nnz == 12
    F = 256
    gen = cuda.Generator(device="cuda").manual_seed(0)
    X = cuda.randn((shape[1], F), device="cuda", generator=gen)
    X[:, 64:128:2] = X[:, 64:128:2].abs() * 1e-3  # q
    X[:, 65:128:2] = X[:, 65:128:2] * 15 + 270  # t
    X[:, 200] = 1.0
    Y = csr.apply(X)
    m = csr_array((d, i, p), shape=shape)
    for col in (0, 77, 255):
        assert_same_values(Y[:, col].cpu().numpy(), m @ X[:, col].cpu().numpy(), f"column {col}")
    w = d.reshape(-1, 12)
    rowsum = np.zeros(shape[0], dtype=np.float32)
    for k in range(12):
        rowsum = rowsum + w[:, k]
    assert_same_values(Y[:, 200].cpu().numpy(), rowsum, "constant field = sequential row sum")
    CL, CH, MK = _cabi.COL_CLIP_LO, _cabi.COL_CLIP_HI, _cabi.COL_MASK
    segs = [(_cabi.EPI_UV2DDFF, 0, 64, 0), (_cabi.EPI_QT2QTR, 64, 64, 64), (_cabi.EPI_PLAIN, 128, 128, 160)]
    cols = [(0, 0, 0, MK)] * 64 + [(0, 0, 0, 0), (0, 0, 0, 0), (0, 100, 85000.0, CL | CH | MK)] * 32 + [(-1.0, 1.0, 0, CL | CH | MK)] * 128
    epi = Epilogue(segs, cols)
    mask = (cuda.rand(shape[0], device="cuda", generator=gen) < 0.3).to(cuda.uint8)
    fused = epi.apply_fused(csr, X, row_mask=mask)
    unfused = epi.apply(Y, row_mask=mask)
    assert cuda.equal(fused.view(cuda.int32), unfused.view(cuda.int32))  # bitwise, NaNs included
    assert bool(cuda.isnan(fused[mask.bool()][:, :64]).all()) and not bool(cuda.isnan(fused[~mask.bool()][:, 160:288]).any())


def test_config5_full_size_lam_masks(cuda):
    """1000 x 1000 LAM at 2 km inside O1280: global_on_lam_mask, thinning_mask, cutout_mask."""
    from anemoi_transform_b200 import spatial
    from anemoi_transform_b200.constants import R_earth_km

    lam = syn.rotated_lam(1000, 1000, 0.018, 60.0, 10.0)
    glob = syn.octahedral(1280)
    rng = np.random.default_rng(3)
    L = np.array(osp.latlon_to_xyz(*lam)).T
    G = np.array(osp.latlon_to_xyz(*glob)).T
    lam_tree = cKDTree(L)

    # global_on_lam_mask with an explicit radius: membership is "some LAM point within r" — checked
    # exactly (cKDTree on the LAM) for every selected point and for a sample of the others
    r_km = 12.0
    gol = spatial.global_on_lam_mask(*lam, *glob, distance_km=r_km)
    assert gol.dtype == np.int64 and np.all(np.diff(gol) > 0) and gol.size > 10_000
    r = r_km / R_earth_km
    d_in = lam_tree.query(G[gol], k=1)[0]
    assert np.all(d_in * d_in <= r * r)
    others = np.setdiff1d(rng.choice(G.shape[0], 300_000, replace=False), gol)
    near = others[np.abs(glob[0][others] - 60.0) < 15.0]
    d_out = lam_tree.query(G[near], k=1)[0]
    assert near.size > 1000 and np.all(d_out * d_out > r * r)

    # thinning_mask: the returned LAM index is at the nearest-neighbour distance (bitwise) for a sample
    thin = spatial.thinning_mask(*lam, *glob)
    crop = osp._crop(lam[0], lam[1], glob[0], glob[1], 2.0)
    assert thin.shape == (int(crop.sum()),) and thin.min() >= 0 and thin.max() < L.shape[0]
    cropped = np.nonzero(crop)[0]
    pick = rng.choice(cropped.size, 20_000, replace=False)
    want_d = lam_tree.query(G[cropped[pick]], k=1)[0]
    got_d = np.sqrt(((L[thin[pick]] - G[cropped[pick]]) ** 2).sum(axis=1))
    assert np.array_equal(got_d, want_d)

    # cutout_mask with numeric distances depends, per global point, on the LAM only: the full mask
    # restricted to a subset equals the oracle run on that subset
    kw = dict(min_distance_km=3.0, max_distance_km=400.0)
    full = spatial.cutout_mask(*lam, *glob, **kw)
    assert full.dtype == np.bool_ and full.shape == glob[0].shape
    box = np.nonzero((np.abs(glob[0] - 60.0) < 14.0) & ((glob[1] < 32.0) | (glob[1] > 348.0)))[0]
    sub = np.sort(np.concatenate([rng.choice(box, 30_000, replace=False), rng.choice(G.shape[0], 5_000, replace=False)]))
    want = osp.cutout_mask_vectorised(*lam, glob[0][sub], glob[1][sub], **kw)
    assert np.array_equal(full[sub], want) and 0 < want.sum() < want.size
