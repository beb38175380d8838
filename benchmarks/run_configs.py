#!/usr/bin/env python
"""Every BASELINE.json config on one GPU, next to the reference's CPU path on the same box.

    python benchmarks/run_configs.py [--out gpurun_out/configs.json] [--only 1,2,4,5]

Config 3 (the headline) is bench.py.  For each other config this reports the GPU path
through the drop-in API (host arrays in, host arrays out), the kernel-only figure where it is
informative, the oracle (= the reference's scipy / cKDTree calls) timed on the host, and a
parity verdict computed on the very arrays that were timed.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.sparse import csr_array  # noqa: E402

from anemoi_transform_b200 import _cabi, ekd, spatial  # noqa: E402
from anemoi_transform_b200 import synthetic as syn  # noqa: E402
from anemoi_transform_b200.device import CsrMatrix, DeviceBatch, Epilogue, KnnIndex  # noqa: E402
from anemoi_transform_b200.filters import create_filter_by_name  # noqa: E402
from oracle import spatial as osp  # noqa: E402


def wall(fn, repeat=3):
    best, out = float("inf"), None
    for _ in range(repeat):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


def dev_ms(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])) if a.dtype.kind == "f" else bool(np.array_equal(a, b))


def config1(tmp):
    s_lat, s_lon = syn.regular_latlon(1.0)
    t_lat, t_lon = syn.octahedral(96)
    d, i, p, shape = syn.bilinear_matrix(1.0, t_lat, t_lon)
    syn.save_regrid_npz(tmp / "c1.npz", d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    fields = [syn.synthetic_field("t", shape[1], s, 0.001 if s % 8 == 0 else 0) for s in range(64)]
    fl = ekd.from_source("list-of-dicts", [dict(param="t", levelist=s, values=v, latitudes=s_lat, longitudes=s_lon) for s, v in enumerate(fields)])
    flt = create_filter_by_name("regrid", matrix=str(tmp / "c1.npz"))
    gpu_s, out = wall(lambda: np.stack([f.to_numpy(flatten=True) for f in flt.forward(fl)]))
    m = csr_array((d, i, p), shape=shape)
    cpu_s, ref = wall(lambda: np.stack([m @ f for f in fields]))
    return {"config": "1: regrid filter, 1deg -> O96, 4-nnz bilinear .npz, 64 float32 fields, FieldList in -> numpy out",
            "gpu_fields_per_s": 64 / gpu_s, "cpu_ref_fields_per_s_1thread": 64 / cpu_s, "bit_exact": same(out, ref)}


def config3_dropin(tmp):
    """Config 3's grids through the drop-in filter API with ordinary (pageable) numpy fields:
    FieldList in -> FieldList out -> to_numpy of every field, 512 fields per call."""
    s_lat, s_lon = syn.regular_latlon(0.25)
    t_lat, t_lon = syn.n320_like()
    d, i, p, shape = syn.bilinear_matrix(0.25, t_lat, t_lon)
    syn.save_regrid_npz(tmp / "c3.npz", d, i, p, shape, s_lat, s_lon, t_lat, t_lon)
    n = 512
    rng = np.random.default_rng(0)
    fields = [rng.standard_normal(shape[1], dtype=np.float32) for _ in range(n)]
    fl = ekd.from_source("list-of-dicts", [dict(param="t", levelist=s, values=v, latitudes=s_lat, longitudes=s_lon) for s, v in enumerate(fields)])
    flt = create_filter_by_name("regrid", matrix=str(tmp / "c3.npz"))

    def run():
        return [f.to_numpy(flatten=True) for f in flt.forward(fl)]

    gpu_s, out = wall(run)
    m = csr_array((d, i, p), shape=shape)
    cpu_s, ref = wall(lambda: [m @ f for f in fields[:32]], repeat=1)
    return {"config": "3 (drop-in): regrid filter 0.25deg -> N320-shaped, 512 pageable float32 fields, FieldList in -> to_numpy of every output",
            "gpu_fields_per_s": n / gpu_s, "gpu_s": gpu_s, "host_GBps_in_plus_out": 4 * n * (shape[0] + shape[1]) / gpu_s / 1e9,
            "cpu_ref_fields_per_s_1thread": 32 / cpu_s, "bit_exact_32_fields": all(same(a, b) for a, b in zip(out[:32], ref))}


def config2():
    src, tgt = syn.regular_latlon(0.25), syn.n320_like()
    gpu_s, (idx, dist, ties) = wall(lambda: spatial.nearest_grid_points(*src, *tgt, _return_ties=True))
    cpu_s, (iref, dref) = wall(lambda: osp.nearest_grid_points(*src, *tgt, return_distances=True), repeat=1)
    sx = spatial.latlon_to_xyz(*src)
    q = tuple(torch.from_numpy(a).cuda() for a in spatial.latlon_to_xyz(*tgt))
    build_s, knn = wall(lambda: KnnIndex(sx))
    q_ms = dev_ms(lambda: knn.query(q, k=1))
    differ = idx != iref
    lam = syn.rotated_lam(400, 400, 0.05, 50.0, 10.0)
    cm_gpu, mask = wall(lambda: spatial.cutout_mask(*lam, *tgt, min_distance_km=30.0))
    cm_cpu, mref = wall(lambda: osp.cutout_mask_vectorised(*lam, *tgt, min_distance_km=30.0), repeat=1)
    return {"config": "2: nearest_grid_points N320-shaped (542,080) vs 0.25deg (1,038,240), k=1; cutout_mask 400x400 LAM in N320",
            "nearest_grid_points_gpu_s": gpu_s, "nearest_grid_points_cpu_s": cpu_s, "kernel_queries_per_s": len(idx) / (q_ms * 1e-3),
            "knn_build_s": build_s, "distances_bitwise_equal": same(dist, dref), "indices_differing": int(differ.sum()),
            "indices_differing_untied": int((differ & (ties == 0)).sum()), "tie_flagged": int((ties != 0).sum()),
            "cutout_gpu_s": cm_gpu, "cutout_cpu_vectorised_s": cm_cpu, "cutout_mask_equal": same(mask, mref)}


def config4():
    s, t = syn.octahedral(1280), syn.n320_like()
    sx, tx = spatial.latlon_to_xyz(*s), spatial.latlon_to_xyz(*t)
    knn = KnnIndex(sx)
    idx, dist, _ = knn.query(tuple(torch.from_numpy(a).cuda() for a in tx), k=12)
    d, i, p, shape = syn.knn_matrix(idx.cpu().numpy(), dist.cpu().numpy(), sx[0].size)
    del knn
    csr = CsrMatrix(d, i, p, shape)
    F = 1024
    gen = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((shape[1], F), device="cuda", generator=gen)
    X[:, 256:512:2] = X[:, 256:512:2].abs() * 1e-3
    X[:, 257:512:2] = X[:, 257:512:2] * 15 + 270
    CL, CH, MK = _cabi.COL_CLIP_LO, _cabi.COL_CLIP_HI, _cabi.COL_MASK
    segs = [(_cabi.EPI_UV2DDFF, 0, 256, 0), (_cabi.EPI_QT2QTR, 256, 256, 256), (_cabi.EPI_PLAIN, 512, 512, 640)]
    cols = [(0, 0, 0, MK)] * 256 + [(0, 0, 0, 0), (0, 0, 0, 0), (0, 100, 85000.0, CL | CH | MK)] * 128 + [(200.0, 320.0, 0, CL | CH | MK)] * 512
    epi = Epilogue(segs, cols)
    mask = (torch.rand(shape[0], device="cuda", generator=gen) < 0.3).to(torch.uint8)
    Y = torch.empty((shape[0], F), device="cuda")
    Yf = torch.empty((shape[0], 1152), device="cuda")
    nref = int(np.unique(i).size)
    plain_ms = dev_ms(lambda: csr.apply(X, out=Y))
    fused_ms = dev_ms(lambda: epi.apply_fused(csr, X, out=Yf, row_mask=mask))
    unfused_ms = dev_ms(lambda: epi.apply(csr.apply(X, out=Y), out=Yf, row_mask=mask))
    alg_plain = 4 * F * (nref + shape[0]) + 8 * d.size + 4 * (shape[0] + 1)
    alg_fused = 4 * F * nref + 4 * 1152 * shape[0] + 8 * d.size + 4 * (shape[0] + 1) + shape[0]
    m = csr_array((d, i, p), shape=shape)
    xs = X[:, :8].cpu().numpy()
    cpu_s, ref = wall(lambda: np.stack([m @ np.ascontiguousarray(xs[:, f]) for f in range(8)], axis=1), repeat=1)
    fused_host = epi.apply_fused(csr, X, out=Yf, row_mask=mask).cpu().numpy()
    unfused_host = epi.apply(csr.apply(X, out=Y), out=Yf, row_mask=mask).cpu().numpy()
    return {"config": "4: regrid O1280 (6,599,680) -> N320-shaped, 12 nnz/row, 1024 float32 fields; + fused uv_to_ddff / q_to_r(all) / clip / mask epilogue (1152 output fields)",
            "plain_ms": plain_ms, "plain_fields_per_s": F / (plain_ms * 1e-3), "plain_algorithmic_GBps": alg_plain / plain_ms / 1e6,
            "fused_ms": fused_ms, "fused_algorithmic_GBps": alg_fused / fused_ms / 1e6, "unfused_spmm_plus_pointwise_ms": unfused_ms,
            "cpu_ref_fields_per_s_1thread": 8 / cpu_s, "plain_bit_exact_8_fields": same(csr.apply(X, out=Y)[:, :8].cpu().numpy(), ref),
            "fused_equals_unfused_bitwise": same(fused_host, unfused_host), "n_src_referenced": nref}


def config3_f64():
    """Config 3's matrix in the dtypes the operational chain produces: MIR writes float64 weights,
    GRIB decodes to float64 values; numpy promotes the mixed cases to float64 results."""
    t_lat, t_lon = syn.n320_like()
    d, i, p, shape = syn.bilinear_matrix(0.25, t_lat, t_lon)
    nref = int(np.unique(i).size)
    m64 = csr_array((d.astype(np.float64), i, p), shape=shape)
    out = {"config": "3 (dtype variants): regrid 0.25deg -> N320-shaped, 4-nnz bilinear, float64 results, kernel-only", "n_src_referenced": nref}
    gen = torch.Generator(device="cuda").manual_seed(0)
    for wname, wdt, xname, xdt, F in (("f64", np.float64, "f64", torch.float64, 1560), ("f64", np.float64, "f32", torch.float32, 1560), ("f32", np.float32, "f64", torch.float64, 1560)):
        csr = CsrMatrix(d.astype(wdt), i, p, shape)
        X = torch.randn((shape[1], F), device="cuda", dtype=xdt, generator=gen) * 15 + 280
        Y = torch.empty((shape[0], F), device="cuda", dtype=torch.float64)
        xs, ws = X.element_size(), np.dtype(wdt).itemsize
        alg = xs * F * nref + 8 * F * shape[0] + (4 + ws) * d.size + 4 * (shape[0] + 1)
        wide = dev_ms(lambda: csr.apply(X, out=Y))
        y_wide = Y[:, :: F // 3].cpu().numpy()
        scalar = dev_ms(lambda: csr.apply(X, out=Y, variant=0x200), n=3, warm=1)
        y_scalar = Y[:, :: F // 3].cpu().numpy()
        mm = m64 if wdt == np.float64 else csr_array((d, i, p), shape=shape)
        ref = np.stack([mm @ X[:, c].cpu().numpy() for c in range(0, F, F // 3)], axis=1)
        out[f"matrix_{wname}_fields_{xname}"] = {"fields": F, "ms": wide, "fields_per_s": F / (wide * 1e-3), "algorithmic_GBps": alg / wide / 1e6,
                                               "scalar_column_kernel_ms": scalar, "bit_exact_sampled_columns": same(y_wide, ref) and same(y_scalar, ref)}
        del csr, X, Y
        torch.cuda.empty_cache()
    return out


def config5():
    lam = syn.rotated_lam(1000, 1000, 0.018, 60.0, 10.0)
    glob = syn.octahedral(1280)
    out = {"config": "5: 2 km LAM (1000x1000, 0.018deg rotated) inside O1280 (6,599,680): global_on_lam_mask + thinning_mask + cutout_mask"}
    out["global_on_lam_mask_gpu_s"], gm = wall(lambda: spatial.global_on_lam_mask(*lam, *glob), repeat=2)
    out["thinning_mask_gpu_s"], tm = wall(lambda: spatial.thinning_mask(*lam, *glob), repeat=2)
    # the LAM straddles lon 0, so the reference's crop box spans every longitude: 570 k queries,
    # most of them tens of degrees from the LAM (the far-query path)
    out["cutout_mask_gpu_s"], cm = wall(lambda: spatial.cutout_mask(*lam, *glob), repeat=2)
    out["global_on_lam_mask_cpu_s"], gref = wall(lambda: osp.global_on_lam_mask(*lam, *glob), repeat=1)
    out["thinning_mask_cpu_s"], tref = wall(lambda: osp.thinning_mask(*lam, *glob), repeat=1)
    out["cutout_mask_cpu_vectorised_s"], cref = wall(lambda: osp.cutout_mask_vectorised(*lam, *glob), repeat=1)
    differ = np.nonzero(tm != tref)[0]
    lp = np.array(spatial.latlon_to_xyz(*lam)).T
    crop = osp._crop(lam[0], lam[1], glob[0], glob[1], 2.0)
    gp = np.array(spatial.latlon_to_xyz(glob[0][crop], glob[1][crop])).T
    d_mine = ((lp[tm[differ]] - gp[differ]) ** 2).sum(axis=1)
    d_ref = ((lp[tref[differ]] - gp[differ]) ** 2).sum(axis=1)
    out.update(thinning_indices_differing=int(differ.size), thinning_differences_are_exact_ties=bool(np.array_equal(d_mine, d_ref)))
    out.update(global_on_lam_mask_equal=same(gm, gref), thinning_mask_equal=same(tm, tref), cutout_mask_equal=same(cm, cref),
               n_global_on_lam=int(gm.size), n_thinning_queries=int(tm.size), n_cutout_dropped=int((~cm).sum()))
    return out


def main():
    import tempfile

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--only", default="1,2,3d,3f64,4,5")
    a = ap.parse_args()
    _cabi.load(check_device=True)
    results = {"host_cpus": os.cpu_count(), "gpu": torch.cuda.get_device_name(0)}
    with tempfile.TemporaryDirectory() as tmp:
        for key, fn in (("1", lambda: config1(Path(tmp))), ("2", config2), ("3d", lambda: config3_dropin(Path(tmp))), ("3f64", config3_f64), ("4", config4), ("5", config5)):
            if key in a.only.split(","):
                t0 = time.perf_counter()
                results[f"config{key}"] = fn()
                results[f"config{key}"]["wall_s_total"] = time.perf_counter() - t0
                print(json.dumps({f"config{key}": results[f"config{key}"]}), flush=True)
                torch.cuda.empty_cache()
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
