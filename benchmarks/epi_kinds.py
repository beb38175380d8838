#!/usr/bin/env python
"""Every pointwise kind on a resident batch (542,080 points x 1024 float32 input columns):
ms, algorithmic GB/s and the fraction of the measured HBM peak.  CUDA events, warm, best of 5.

    python benchmarks/epi_kinds.py [--json out.json]
    AT_UNDER_NCU=1 ncu --set full -k regex:pointwise_kernel … python benchmarks/epi_kinds.py
        one launch per (kind, flags), in the order of the table — numbers printed under ncu
        are not bench values
"""

from __future__ import annotations

import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200"), str(REPO / "benchmarks")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
from run_configs import dev_ms  # noqa: E402

from anemoi_transform_b200 import _cabi  # noqa: E402
from anemoi_transform_b200.device import Epilogue  # noqa: E402


def main():
    _cabi.load(check_device=True)
    peak = 6537.6
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        peak = float(json.loads(f.read_text())["hbm_gbs"])
    n, F = 542080, 1024
    gen = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((n, F), device="cuda", generator=gen)
    Xuv = X * 8.0
    Xqt = X.clone()
    Xqt[:, 0::2] = Xqt[:, 0::2].abs() * 1e-3
    Xqt[:, 1::2] = Xqt[:, 1::2] * 15 + 270
    Xrt = Xqt.clone()
    Xrt[:, 0::2] = (X[:, 0::2].abs() * 40).clamp(max=100.0)
    Xdd = X.clone()
    Xdd[:, 0::2] = Xdd[:, 0::2].abs() * 10
    Xdd[:, 1::2] = (Xdd[:, 1::2] * 100) % 360
    mask = (torch.rand(n, device="cuda", generator=gen) < 0.3).to(torch.uint8)
    CL, CH, MK = _cabi.COL_CLIP_LO, _cabi.COL_CLIP_HI, _cabi.COL_MASK
    rows = []
    kinds = (
        ("plain", _cabi.EPI_PLAIN, F, X),
        ("uv2ddff", _cabi.EPI_UV2DDFF, F, Xuv),
        ("ddff2uv", _cabi.EPI_DDFF2UV, F, Xdd),
        ("qt2r", _cabi.EPI_QT2R, F // 2, Xqt),
        ("qt2qtr", _cabi.EPI_QT2QTR, F * 3 // 2, Xqt),
        ("rt2q", _cabi.EPI_RT2Q, F // 2, Xrt),
        ("rt2d", _cabi.EPI_RT2D, F // 2, Xrt),
        ("cossin", _cabi.EPI_COSSIN, F * 2, X),
        ("atan2", _cabi.EPI_ATAN2, F // 2, X),
    )
    for name, kind, nout, x in kinds:
        for fl_name, fl in (("noflags", 0), ("clip+mask", CL | CH | MK)):
            epi = Epilogue([(kind, 0, F, 0, 1.0, 0.0)], [(-1e30, 1e30, 85000.0, fl)] * nout)
            Y = torch.empty((n, (nout + 3) // 4 * 4), device="cuda")
            if os.environ.get("AT_UNDER_NCU") and fl_name != "noflags" and name not in ("plain", "uv2ddff"):
                continue  # the capture takes ~10 s per launch: flags on two kinds are enough
            if os.environ.get("AT_UNDER_NCU"):
                epi.apply(x, out=Y, row_mask=mask)
                torch.cuda.synchronize()
                print(f"launched {name} {fl_name}", flush=True)
                continue
            ms = dev_ms(lambda: epi.apply(x, out=Y, row_mask=mask))
            gbs = 4 * n * (F + nout) / ms / 1e6
            rows.append({"kind": name, "flags": fl_name, "ms": ms, "algorithmic_GBps": gbs, "frac_of_measured_peak": gbs / peak})
            print(f"{name:8s} {fl_name:10s} {ms:.3f} ms  {gbs:.0f} GB/s  {gbs / peak:.2f}", flush=True)
            del Y
    if "--json" in sys.argv:
        Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
