#!/usr/bin/env python
"""grib_unpack_kernel alone: packed values resident in HBM -> columns of the point-major batch.
Algorithmic bytes = F * (ceil(n * nbits / 8) + elem * n); CUDA events, warm, best of 5.
Under ncu (AT_UNDER_NCU=1): one launch per case."""
import ctypes
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

from anemoi_transform_b200 import _cabi, grib  # noqa: E402
from anemoi_transform_b200.device import _ptr, stream_ptr  # noqa: E402
from oracle import grib as ogrib  # noqa: E402

_cabi.load(check_device=True)
UNDER_NCU = bool(os.environ.get("AT_UNDER_NCU"))
n_points, n_fields = 1_038_240, 256
peak = 6537.6
try:
    peak = float(json.loads((REPO / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
rng = np.random.default_rng(0)
out = {"n_points": n_points, "n_fields": n_fields, "hbm_peak_GBps": peak, "cases": []}
cases = [(16, "float64"), (16, "float32")] if UNDER_NCU else [(nb, dt) for nb in (8, 12, 16, 24, 11) for dt in ("float64", "float32")]
for nbits, dt in cases:
    distinct = [ogrib.encode_grib2(rng.normal(280.0, 15.0, n_points), nbits, 0) for _ in range(4)]
    infos = (_cabi.GribInfo * n_fields)(*[grib.scan(distinct[k % 4]) for k in range(n_fields)])
    blob, offs = bytearray(), []
    for k in range(n_fields):
        i = infos[k]
        while len(blob) % 256:
            blob.append(0)
        offs.append(len(blob))
        blob += distinct[k % 4][i.data_offset : i.data_offset + i.data_length]
    d_blob = torch.frombuffer(blob, dtype=torch.uint8).cuda()
    tdt = torch.float64 if dt == "float64" else torch.float32
    y = torch.empty((n_points, n_fields), dtype=tdt, device="cuda")
    offsets = (ctypes.c_int64 * n_fields)(*offs)
    code = _cabi.AT_F64 if dt == "float64" else _cabi.AT_F32

    def run():
        _cabi.call("at_grib_unpack", _ptr(d_blob), offsets, infos, n_fields, n_points, code, _ptr(y), n_fields, stream_ptr())

    run()
    if UNDER_NCU:
        continue
    best = float("inf")
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    # the entry point uploads its parameter table and frees it synchronously: time the kernel
    # through a longer queue as well
    alg = n_fields * ((n_points * nbits + 7) // 8 + y.element_size() * n_points)
    out["cases"].append({"bits": nbits, "out": dt, "ms_call": best, "algorithmic_bytes": alg, "GBps_call": alg / best / 1e6, "frac_of_peak_call": alg / best / 1e6 / peak})
    want = ogrib.decode(distinct[1])
    got = y[:, 1].cpu().numpy()
    assert np.array_equal(got, want if dt == "float64" else want.astype(np.float32))
    del d_blob, y
print(json.dumps(out, indent=1))
