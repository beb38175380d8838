"""The SpMM oracle: restatements (numpy, plain C) against live scipy, and scipy's call
against the golden outputs of the imported reference RegridFilter."""

import numpy as np
import pytest
from conftest import assert_same_values
from scipy.sparse import csr_array

from anemoi_transform_b200 import synthetic as syn
from oracle import spmm


def _matrix32():
    t_lat, t_lon = syn.octahedral(16)
    return syn.bilinear_matrix(5.0, t_lat, t_lon)


def test_restatements_match_scipy_bitwise():
    d, i, p, shape = _matrix32()
    m = csr_array((d, i, p), shape=shape)
    fields = np.stack([syn.synthetic_field("t", shape[1], s, 0.01) for s in range(5)])
    fields[0, 3] = np.inf
    want = np.stack([m @ x for x in fields])
    assert_same_values(np.stack([spmm.csr_matvec_sequential(p, i, d, x) for x in fields]), want, "numpy restatement")
    assert_same_values(spmm.c_regrid_fields_f32(p, i, d, fields), want, "C restatement")
    assert_same_values(spmm.c_regrid_fields_f32(p, i, d, fields, n_threads=1), want, "C restatement, 1 thread")


def test_sequential_unfused_semantics():
    """The facts the GPU kernel mirrors: order matters, explicit zeros propagate NaN / inf,
    empty rows give 0, result dtype is numpy's result_type."""
    indptr = np.array([0, 3, 3, 5, 6])
    indices = np.array([0, 1, 2, 0, 1, 2])
    data = np.array([1e8, 1.0, -1e8, 0.0, 0.0, 0.5], dtype=np.float32)
    x = np.array([1.0, 1.0, 1.0], dtype=np.float32)
    m = csr_array((data, indices, indptr), shape=(4, 3))
    y = m @ x
    assert y[0] == np.float32(np.float32(np.float32(1e8) + np.float32(1.0)) - np.float32(1e8))  # sequential
    assert y[1] == 0.0  # empty row
    xn = np.array([np.nan, np.inf, 2.0], dtype=np.float32)
    yn = m @ xn
    assert np.isnan(yn[2])  # 0*nan + 0*inf
    assert yn[3] == 1.0
    assert_same_values(spmm.csr_matvec_sequential(indptr, indices, data, xn), yn)
    assert (csr_array((data.astype(np.float64), indices, indptr), shape=(4, 3)) @ x).dtype == np.float64
    assert spmm.csr_matvec_sequential(indptr, indices, data.astype(np.float64), x).dtype == np.float64


@pytest.mark.parametrize("mat,fld", [("m32", "fields32"), ("m32", "fields64"), ("m64", "fields32"), ("m64", "fields64")])
def test_scipy_call_reproduces_reference_filter_output(golden_regrid, mat, fld):
    g = golden_regrid
    m = csr_array((g[f"{mat}_data"], g[f"{mat}_indices"], g[f"{mat}_indptr"]), shape=tuple(g[f"{mat}_shape"]))
    want = g[f"y_{mat}_{'f32' if fld == 'fields32' else 'f64'}"]
    got = np.stack(spmm.regrid_fields(m, list(g[fld])))
    assert_same_values(got, want, f"{mat} @ {fld}")
    rest = np.stack([spmm.csr_matvec_sequential(m.indptr, m.indices, m.data, x) for x in g[fld]])
    assert_same_values(rest, want, f"restatement {mat} @ {fld}")


def test_gather_goldens_are_plain_indexing(golden_regrid):
    g = golden_regrid
    assert_same_values(g["fields32"][:, g["mask_idx"]], g["y_mask_f32"])
    assert np.array_equal(g["y_mask_lat"], g["s_lat"][g["mask_idx"]])
    from oracle import spatial

    idx = spatial.nearest_grid_points(g["s_lat"], g["s_lon"], g["t_lat"], g["t_lon"])
    assert_same_values(g["fields32"][:, idx], g["y_nearest_f32"])
