#!/usr/bin/env python
"""spmm_f32_kernel on the per-GPU shard of the strong-scaling runs (config 3 sharded over 8 GPUs:
392 fields): ms per launch for the kernel-shape variants of at_spmm."""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "anemoi-transform_b200"), str(REPO / "benchmarks")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
from run_configs import dev_ms  # noqa: E402

from anemoi_transform_b200 import _cabi  # noqa: E402
from anemoi_transform_b200 import synthetic as syn  # noqa: E402
from anemoi_transform_b200.device import CsrMatrix  # noqa: E402

_cabi.load(check_device=True)
t_lat, t_lon = syn.n320_like()
d, i, p, shape = syn.bilinear_matrix(0.25, t_lat, t_lon)
csr = CsrMatrix(d, i, p, shape)
out = {}
for F in (392, 780, 1560, 3120):
    X = torch.randn((shape[1], F), device="cuda")
    Y = torch.empty((shape[0], F), device="cuda")
    row = {}
    for name, variant in (("default", 0), ("vpl1", 1), ("vpl2", 2), ("vpl4", 4), ("vpl4_rpw4", 4 | (4 << 4)), ("vpl2_rpw4", 2 | (4 << 4)), ("vpl1_rpw4", 1 | (4 << 4)), ("vpl4_nobulk", 4 | (1 << 8))):
        try:
            row[name] = dev_ms(lambda: csr.apply(X, out=Y, variant=variant), n=10, warm=3)
        except Exception as e:
            row[name] = str(e)[:60]
    out[F] = row
    ideal = 2.883 * F / 3120
    print(F, f"ideal {ideal:.3f} ms", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}, flush=True)
    del X, Y
print(json.dumps(out))
