"""SpMM on batches whose last column tile is ragged: when that tile fills half a warp or less,
spmm_f32_kernel packs several target rows into the warp (every lane walks the
CSR entries of its own row).  Results must stay scipy's bit for bit (regrid.py:310 `matrix @ x`
→ scipy csr_matvec: sequential, unfused, storage order) for every dtype combination, on uniform
4- and 12-nonzero rows and on ragged rows with empty ones.
"""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _matrix(rng, n_tgt, n_src, nnz_per_row, wdtype):
    lengths = np.full(n_tgt, nnz_per_row) if nnz_per_row else rng.integers(0, 10, n_tgt)
    indptr = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
    indices = np.concatenate([np.sort(rng.choice(n_src, k, replace=False)) for k in lengths] + [np.empty(0, np.int64)]).astype(np.int32)
    data = rng.random(indices.size).astype(wdtype)
    data[rng.random(data.size) < 0.02] = 0.0  # explicit zeros stay: 0 * nan = nan
    return data, indices, indptr


# field counts: 4 / 12 / 40 fill 1 / 3 / 10 lanes of the only tile; 260 and 780 leave a ragged last
# tile behind full ones (float32: 65 = 64 + 1 and 195 = 3 * 64 + 3 units; float64: 130 = 2 * 64 + 2)
@pytest.mark.parametrize("n_fields", [4, 12, 40, 260, 780])
@pytest.mark.parametrize("nnz_per_row", [4, 12, 0])
@pytest.mark.parametrize("wdtype,xdtype", [(np.float32, np.float32), (np.float64, np.float32), (np.float32, np.float64), (np.float64, np.float64)])
def test_ragged_last_tile_is_bit_exact(cuda, n_fields, nnz_per_row, wdtype, xdtype):
    import torch
    from scipy.sparse import csr_array

    from anemoi_transform_b200.device import CsrMatrix

    rng = np.random.default_rng(n_fields * 100 + nnz_per_row)
    n_tgt, n_src = 301, 407
    data, indices, indptr = _matrix(rng, n_tgt, n_src, nnz_per_row, wdtype)
    x = rng.standard_normal((n_src, n_fields)).astype(xdtype)
    x[rng.random(x.shape) < 0.002] = np.nan
    m = csr_array((data, indices, indptr), shape=(n_tgt, n_src))
    csr = CsrMatrix(data, indices, indptr, (n_tgt, n_src))
    got = csr.apply(torch.from_numpy(x).cuda(), n_fields=n_fields).cpu().numpy()[:, :n_fields]
    cols = sorted({0, 1, n_fields // 2, n_fields - 2, n_fields - 1})
    want = np.stack([m @ np.ascontiguousarray(x[:, c]) for c in cols], axis=1)
    assert got.dtype == want.dtype
    g = got[:, cols]
    assert np.array_equal(np.isnan(g), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(g[ok], want[ok]) and np.array_equal(np.signbit(g[ok]), np.signbit(want[ok]))
