#!/usr/bin/env python
"""One launch of every hot kernel at BASELINE sizes — the program `ncu --set full` captures.

    AT_UNDER_NCU=1 ncu --set full --clock-control none --import-source on \\
        -k regex:'spmm_|pointwise_|gather_|transpose_|knn_' -o gpurun_out/prof_kernels python benchmarks/ncu_kernels.py
    python benchmarks/ncu_kernels.py      # the same launches timed with CUDA events (warm, best of 5)

Prints, per kernel, the algorithmic bytes of the launch so the capture's
dram__bytes and duration can be set against them.  Numbers printed under ncu are not bench values.
"""

from __future__ import annotations

import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from anemoi_transform_b200 import _cabi, spatial  # noqa: E402
from anemoi_transform_b200 import synthetic as syn  # noqa: E402
from anemoi_transform_b200.device import CsrMatrix, Epilogue, KnnIndex, gather_rows, transpose  # noqa: E402


UNDER_NCU = bool(os.environ.get("AT_UNDER_NCU"))
MS = {}


def run(name, fn):
    """Under ncu: one launch.  Otherwise: warm-up + best of 5, CUDA events."""
    if UNDER_NCU:
        fn()
        return
    fn()
    best = float("inf")
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    MS[name] = best


def main():
    _cabi.load(check_device=True)
    out = {}
    gen = torch.Generator(device="cuda").manual_seed(0)
    t_lat, t_lon = syn.n320_like()
    d, i, p, shape = syn.bilinear_matrix(0.25, t_lat, t_lon)
    nref = int(np.unique(i).size)
    n_tgt, n_src = shape

    # spmm_f32_kernel, config 3 at 1560 fields (half the headline batch: same per-field bytes)
    F = 1560
    csr = CsrMatrix(d, i, p, shape)
    X = torch.randn((n_src, F), device="cuda", generator=gen)
    # physical value ranges for the columns the epilogue programs below convert: (u, v) in m/s,
    # (q, t) in kg/kg and K — a temperature of N(0, 1) K sends every humidity conversion down the
    # special-value path and measures that instead of the kernel
    X[:, 0:520] *= 8.0
    X[:, 520:1040:2] = X[:, 520:1040:2].abs() * 1e-3
    X[:, 521:1040:2] = X[:, 521:1040:2] * 15.0 + 270.0
    Y = torch.empty((n_tgt, F), device="cuda")
    run("spmm_f32_kernel", lambda: csr.apply(X, out=Y))
    out["spmm_f32_kernel"] = 4 * F * (nref + n_tgt) + 8 * d.size + 4 * (n_tgt + 1)

    # pointwise_kernel<float>: uv_to_ddff + q_to_r(all) + clip + mask on the regridded batch
    CL, CH, MK = _cabi.COL_CLIP_LO, _cabi.COL_CLIP_HI, _cabi.COL_MASK
    segs = [(_cabi.EPI_UV2DDFF, 0, 520, 0), (_cabi.EPI_QT2QTR, 520, 520, 520), (_cabi.EPI_PLAIN, 1040, 520, 1300)]
    cols = [(0, 0, 0, MK)] * 520 + [(0, 0, 0, 0), (0, 0, 0, 0), (0, 100, 85000.0, CL | CH | MK)] * 260 + [(200.0, 320.0, 0, CL | CH | MK)] * 520
    epi = Epilogue(segs, cols)
    mask = (torch.rand(n_tgt, device="cuda", generator=gen) < 0.3).to(torch.uint8)
    Yp = torch.empty((n_tgt, 1820), device="cuda")
    run("pointwise_kernel<float>", lambda: epi.apply(Y, out=Yp, row_mask=mask))
    out["pointwise_kernel<float>"] = 4 * (1560 + 1820) * n_tgt + n_tgt

    # spmm_fused_kernel on the same program
    run("spmm_fused_kernel", lambda: epi.apply_fused(csr, X, out=Yp, row_mask=mask))
    out["spmm_fused_kernel"] = 4 * F * nref + 4 * 1820 * n_tgt + 8 * d.size + 4 * (n_tgt + 1) + n_tgt

    # gather_rows_kernel: nearest-neighbour regrid of the same batch
    sx = spatial.latlon_to_xyz(*syn.regular_latlon(0.25))
    tq = tuple(torch.from_numpy(a).cuda() for a in spatial.latlon_to_xyz(t_lat, t_lon))
    knn = KnnIndex(sx)
    idx, _, _ = knn.query(tq, k=1)
    run("knn_query_kernel", lambda: knn.query(tq, k=1))
    out["knn_query_kernel"] = 24 * n_tgt + 24 * n_src + 16 * n_tgt
    idx0 = idx[:, 0].contiguous()
    run("gather_rows_kernel", lambda: gather_rows(X, idx0, out=Y))
    out["gather_rows_kernel"] = 4 * F * (int(torch.unique(idx).numel()) + n_tgt) + 8 * n_tgt

    # transpose_kernel: pack 256 field-major fields point-major
    fm = torch.randn((256, n_src), device="cuda", generator=gen)
    fmt = torch.empty((n_src, 256), device="cuda")
    run("transpose_kernel<float>", lambda: transpose(fm, out=fmt))
    out["transpose_kernel<float>"] = 2 * 4 * 256 * n_src
    del X, Y, Yp, fm, fmt, csr
    torch.cuda.empty_cache()

    # spmm_f64_kernel: float64 matrix, float64 and float32 fields
    F = 780
    csr64 = CsrMatrix(d.astype(np.float64), i, p, shape)
    for name, dt in (("double,double", torch.float64), ("double,float", torch.float32)):
        X = torch.randn((n_src, F), device="cuda", dtype=dt, generator=gen)
        Y = torch.empty((n_tgt, F), device="cuda", dtype=torch.float64)
        run(f"spmm_f64_kernel<{name}>", lambda: csr64.apply(X, out=Y))
        out[f"spmm_f64_kernel<{name}>"] = X.element_size() * F * nref + 8 * F * n_tgt + 12 * d.size + 4 * (n_tgt + 1)
        del X, Y
    torch.cuda.synchronize()
    print(json.dumps({"algorithmic_bytes_per_launch": out, "ms": MS, "algorithmic_GBps": {k: out[k] / v / 1e6 for k, v in MS.items()}}))


if __name__ == "__main__":
    main()
