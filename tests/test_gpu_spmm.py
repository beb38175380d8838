"""SpMM parity through the C-ABI: bit-exact against the oracle (scipy's csr_matvec)."""

import numpy as np
import pytest
from conftest import assert_same_values
from scipy.sparse import csr_array

from anemoi_transform_b200 import synthetic as syn
from oracle import spmm as ospmm

pytestmark = pytest.mark.gpu

# variant = vpl | rows_per_warp << 4 | no_bulk << 8 | force_general << 9
VARIANTS = [0, 1, 2, 4, 0x11, 0x42, 0x84, 0x101, 0x102, 0x201, 0x202, 0x204]


def _apply(cuda, csr, fields, variant=0):
    """fields: numpy [F, n_src] → numpy [F, n_tgt] through pack / at_spmm / unpack."""
    from anemoi_transform_b200.device import DeviceBatch

    batch = DeviceBatch.from_host_fields(list(fields))
    y = csr.apply(batch.data, n_fields=batch.n_fields, variant=variant)
    return DeviceBatch(y, batch.n_fields).to_host_fields()


@pytest.fixture(scope="module")
def config1(cuda):
    """BASELINE config 1: 1° (360x181) → O96, 4-point bilinear, 64 float32 fields."""
    from anemoi_transform_b200.device import CsrMatrix

    t_lat, t_lon = syn.octahedral(96)
    d, i, p, shape = syn.bilinear_matrix(1.0, t_lat, t_lon)
    fields = np.stack([syn.synthetic_field("t", shape[1], s, 0.001 if s % 8 == 0 else 0.0) for s in range(64)])
    fields[3, 100] = np.inf
    want = np.stack([csr_array((d, i, p), shape=shape) @ x for x in fields])
    return CsrMatrix(d, i, p, shape), fields, want


@pytest.mark.parametrize("variant", VARIANTS)
def test_config1_bit_exact_every_kernel_shape(cuda, config1, variant):
    csr, fields, want = config1
    assert csr.uniform_nnz == 4
    assert_same_values(_apply(cuda, csr, fields, variant), want, f"variant {variant:#x}")


@pytest.mark.parametrize("n_fields", [1, 2, 3, 5, 31, 33, 127, 129, 260])
def test_ragged_field_counts(cuda, config1, n_fields):
    csr, fields, want = config1
    reps = -(-n_fields // fields.shape[0])
    f = np.concatenate([fields] * reps)[:n_fields]
    w = np.concatenate([want] * reps)[:n_fields]
    assert_same_values(_apply(cuda, csr, f), w, f"{n_fields} fields")


@pytest.mark.parametrize("mat,fld", [("m32", "fields32"), ("m32", "fields64"), ("m64", "fields32"), ("m64", "fields64")])
def test_golden_reference_outputs_all_dtype_combinations(cuda, golden_regrid, mat, fld):
    """Outputs of the imported reference RegridFilter; m64 is irregular (empty rows, unsorted
    columns, explicit zeros) and float64; the result dtype follows numpy's result_type."""
    from anemoi_transform_b200.device import CsrMatrix

    g = golden_regrid
    csr = CsrMatrix(g[f"{mat}_data"], g[f"{mat}_indices"], g[f"{mat}_indptr"], tuple(g[f"{mat}_shape"]))
    want = g[f"y_{mat}_{'f32' if fld == 'fields32' else 'f64'}"]
    assert_same_values(_apply(cuda, csr, g[fld]), want, f"{mat} @ {fld}")


def test_irregular_float32_matrix_general_kernel(cuda):
    """Rows of 0…40 entries incl. a row longer than the staged segment capacity."""
    from anemoi_transform_b200.device import CsrMatrix

    rng = np.random.default_rng(5)
    n_t, n_s = 1500, 4000
    lens = rng.integers(0, 41, n_t)
    lens[7] = 3000  # spills the shared-memory segment: global-memory path
    lens[8] = 0
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = rng.integers(0, n_s, ptr[-1]).astype(np.int32)
    dat = rng.normal(size=ptr[-1]).astype(np.float32)
    dat[rng.uniform(size=dat.size) < 0.05] = 0.0
    fields = rng.normal(size=(37, n_s)).astype(np.float32)
    fields[2, 17] = np.nan
    m = csr_array((dat, idx, ptr), shape=(n_t, n_s))
    want = np.stack([m @ x for x in fields])
    csr = CsrMatrix(dat, idx, ptr, (n_t, n_s))
    assert csr.uniform_nnz == 0
    for variant in (0, 1, 4, 0x81):
        assert_same_values(_apply(cuda, csr, fields, variant), want, f"irregular, variant {variant:#x}")
    assert_same_values(np.stack([ospmm.csr_matvec_sequential(ptr, idx, dat, x) for x in fields]), want, "oracle restatement")


@pytest.mark.parametrize("wdt,xdt", [(np.float64, np.float64), (np.float64, np.float32), (np.float32, np.float64)])
def test_float64_results_vectorised_and_scalar_kernels(cuda, wdt, xdt):
    """float64 matrix and / or float64 fields (MIR weights are float64, GRIB values decode to
    float64): spmm_f64_kernel (4-field groups) and the scalar-column kernel (variant bit 9)
    against scipy, on the regular config-1 matrix and on an irregular one whose row 7 spills
    the staged segment (global-memory path), for ragged field counts."""
    from anemoi_transform_b200.device import CsrMatrix

    rng = np.random.default_rng(11)
    t_lat, t_lon = syn.octahedral(48)
    d, i, p, shape = syn.bilinear_matrix(1.0, t_lat, t_lon)
    n_t, n_s = 1500, 4000
    lens = rng.integers(0, 41, n_t)
    lens[7], lens[8] = 3000, 0
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = rng.integers(0, n_s, ptr[-1]).astype(np.int32)
    dat = rng.normal(size=ptr[-1])
    dat[rng.uniform(size=dat.size) < 0.05] = 0.0
    for name, (md, mi, mp, mshape) in {"bilinear": (d, i, p, shape), "irregular": (dat, idx, ptr, (n_t, n_s))}.items():
        md = md.astype(wdt)
        m = csr_array((md, mi, mp), shape=mshape)
        csr = CsrMatrix(md, mi, mp, mshape)
        for n_fields in (1, 3, 4, 37, 129, 260):
            fields = (rng.normal(size=(n_fields, mshape[1])) * 10 + 280).astype(xdt)
            fields[0, 17] = np.nan
            fields[n_fields // 2, 100] = np.inf
            want = np.stack([m @ x for x in fields])
            assert want.dtype == np.float64
            for variant in (0, 1, 4, 0x1000, 0x200):
                got = _apply(cuda, csr, fields, variant)
                assert got.dtype == np.float64
                assert_same_values(got, want, f"{name} {wdt.__name__}@{xdt.__name__} F={n_fields} variant {variant:#x}")


def test_empty_and_degenerate(cuda):
    from anemoi_transform_b200.device import CsrMatrix

    csr = CsrMatrix(np.zeros(0, np.float32), np.zeros(0, np.int32), np.zeros(6, np.int32), (5, 9))
    x = cuda.ones((9, 4), device="cuda")
    assert cuda.equal(csr.apply(x), cuda.zeros((5, 4), device="cuda"))  # empty rows → 0
    with pytest.raises(ValueError):
        CsrMatrix(np.ones(1, np.float32), np.array([9], np.int32), np.array([0, 1], np.int32), (1, 9))  # column out of range
    with pytest.raises(ValueError):
        CsrMatrix(np.ones(2, np.float32), np.array([0, 1], np.int32), np.array([0, 2, 1], np.int32), (2, 9))  # indptr not monotone
    with pytest.raises(ValueError):
        csr.apply(cuda.ones((8, 4), device="cuda"))  # wrong number of source points


def test_full_size_config3_properties(cuda):
    """0.25° → N320, 3120 fields: size-independent checks at BASELINE's full size.

    (a) a constant field regrids to the sequential sum of each row's weights (bitwise);
    (b) sampled columns agree bitwise with scipy;
    (c) linearity in exact arithmetic: doubling the input doubles the output bitwise."""
    from anemoi_transform_b200.device import CsrMatrix

    t_lat, t_lon = syn.n320_like()
    d, i, p, shape = syn.bilinear_matrix(0.25, t_lat, t_lon)
    assert shape == (542_080, 1_038_240)
    csr = CsrMatrix(d, i, p, shape)
    n_fields = 3120
    gen = cuda.Generator(device="cuda").manual_seed(3)
    x = cuda.randn((shape[1], n_fields), device="cuda", dtype=cuda.float32, generator=gen)
    x[:, 5] = 1.0
    y = csr.apply(x)
    w = d.reshape(-1, 4)
    rowsum = ((np.float32(0) + w[:, 0]) + w[:, 1] + w[:, 2]) + w[:, 3]
    assert_same_values(y[:, 5].cpu().numpy(), rowsum, "constant field")
    m = csr_array((d, i, p), shape=shape)
    for col in (0, 1234, 3119):
        assert_same_values(y[:, col].cpu().numpy(), m @ x[:, col].cpu().numpy(), f"column {col}")
    y2 = csr.apply(x * 2)
    assert cuda.equal(y2, y * 2)


def test_layout_kernels(cuda):
    """at_transpose, at_gather_cols, at_gather_rows and the chunked batch upload / download."""
    from anemoi_transform_b200.device import DeviceBatch, gather_cols, gather_rows, transpose

    rng = np.random.default_rng(2)
    for dtype in (np.float32, np.float64):
        for rows, cols in ((1, 1), (3, 130), (65, 64), (200, 257)):
            a = rng.normal(size=(rows, cols)).astype(dtype)
            t = transpose(cuda.from_numpy(a).cuda()).cpu().numpy()
            assert np.array_equal(t, a.T)
        # the 16-byte path (leading dimensions that allow aligned float4 / double2 accesses), with
        # ragged edges in both directions taken from views of padded arrays
        for rows, cols, ld_src, ld_dst in ((64, 64, 64, 64), (68, 132, 132, 68), (61, 130, 132, 64), (200, 257, 260, 200), (5, 3, 4, 8), (130, 1000, 1000, 132)):
            big = cuda.from_numpy(rng.normal(size=(rows, ld_src)).astype(dtype)).cuda()
            out = cuda.full((cols, ld_dst), -7.0, dtype=big.dtype, device="cuda")
            transpose(big[:, :cols], out=out[:, :rows])
            assert cuda.equal(out[:, :rows], big[:, :cols].T.contiguous()), (rows, cols, ld_src, ld_dst)
            assert bool((out[:, rows:] == -7.0).all())  # nothing written beyond the requested block
        fields = [rng.normal(size=1000).astype(dtype) for _ in range(150)]
        fields[3][7] = np.nan
        b = DeviceBatch.from_host_fields(fields, chunk=64)
        assert b.data.shape == (1000, 152) and b.n_fields == 150
        assert_same_values(b.to_host_fields(chunk=32), np.stack(fields), f"round trip {dtype.__name__}")
        pick = [149, 0, 3, 3, 77]
        g = gather_cols(b.data, pick).cpu().numpy()
        assert_same_values(np.ascontiguousarray(g[:, :5].T), np.stack([fields[p] for p in pick]), "gather_cols")
        idx = cuda.from_numpy(rng.integers(0, 1000, 333)).cuda()
        r = gather_rows(b.data, idx).cpu().numpy()
        assert_same_values(r[:, :150], np.stack(fields).T[idx.cpu().numpy()], "gather_rows")
    with pytest.raises(IndexError):
        gather_cols(b.data, [152])
    with pytest.raises(IndexError):
        gather_rows(b.data, cuda.tensor([1000], device="cuda"))
    mixed = DeviceBatch.from_host_fields([np.ones(10, np.float32), np.ones(10, np.float64)])
    assert mixed.data.dtype == cuda.float64  # numpy result_type


@pytest.mark.parametrize("seed", range(12))
def test_random_matrices_bit_exact_against_scipy(cuda, seed):
    """Fuzz: random shapes, row lengths 0…60 (duplicate and unsorted columns, explicit zeros,
    int32 / int64 index arrays), float32 / float64 weights and fields, ragged field counts,
    NaN / inf / signed zeros in the fields — every output bit equal to scipy's `csr @ x`."""
    from anemoi_transform_b200.device import CsrMatrix

    rng = np.random.default_rng(1000 + seed)
    n_t, n_s = int(rng.integers(1, 3000)), int(rng.integers(1, 5000))
    max_len = int(rng.choice([1, 4, 12, 60]))
    lens = rng.integers(0, max_len + 1, n_t)
    if seed % 4 == 0:
        lens[:] = max_len  # uniform row length: the bulk-staged kernels
    if seed % 5 == 1 and n_t > 3:
        lens[2] = 2500  # one row longer than the staged segment
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(rng.choice([np.int32, np.int64]))
    idx = rng.integers(0, n_s, int(ptr[-1])).astype(rng.choice([np.int32, np.int64]))
    wdt = rng.choice([np.float32, np.float64])
    xdt = rng.choice([np.float32, np.float64])
    dat = rng.normal(size=int(ptr[-1])).astype(wdt)
    dat[rng.uniform(size=dat.size) < 0.1] = 0.0
    n_fields = int(rng.choice([1, 2, 3, 4, 5, 31, 64, 129, 257]))
    fields = (rng.normal(size=(n_fields, n_s)) * rng.choice([1e-3, 1.0, 1e6])).astype(xdt)
    for special in (np.nan, np.inf, -np.inf, -0.0):
        fields[rng.integers(0, n_fields), rng.integers(0, n_s)] = special
    m = csr_array((dat, idx, ptr), shape=(n_t, n_s))
    with np.errstate(invalid="ignore", over="ignore"):
        want = np.stack([m @ x for x in fields])
    csr = CsrMatrix(dat, idx, ptr, (n_t, n_s))
    for variant in (0, 1, 0x200):
        got = _apply(cuda, csr, fields, variant)
        assert_same_values(got, want, f"seed {seed}: {n_t}x{n_s}, rows <= {max_len}, {np.dtype(wdt).name} @ {np.dtype(xdt).name}, F={n_fields}, variant {variant:#x}")
