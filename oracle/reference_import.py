"""Import the UNMODIFIED reference from /root/reference behind oracle/refstubs.

TEST INFRASTRUCTURE, build container only: /root/reference does not exist on the GPU box,
so nothing that runs there may call this.  Used by oracle/make_golden.py and by the CPU
tests that validate the oracle restatements against the live reference (skipped when the
reference tree is absent).
"""

from __future__ import annotations

import importlib
import os
import sys
from pathlib import Path

REFERENCE_SRC = Path(os.environ.get("ANEMOI_REFERENCE_SRC", "/root/reference/src"))
STUBS = Path(__file__).resolve().parent / "refstubs"
REPO = Path(__file__).resolve().parent.parent

HOT_PATH_MODULES = (
    "anemoi.transform.workflows.pipeline",
    "anemoi.transform.spatial",
    "anemoi.transform.filters.fields.regrid",
    "anemoi.transform.filters.fields.uv_to_ddff",
    "anemoi.transform.filters.fields.q_to_r",
    "anemoi.transform.filters.fields.clipper",
    "anemoi.transform.filters.fields.apply_mask",
    # the remaining pointwise field filters
    "anemoi.transform.filters.fields.rescale",
    "anemoi.transform.filters.fields.lnsp_to_sp",
    "anemoi.transform.filters.fields.impute_nans",
    "anemoi.transform.filters.fields.remove_nans",
    "anemoi.transform.filters.fields.cos_sin_from_rad",
    "anemoi.transform.filters.fields.cos_sin_mean_wave_direction",
    "anemoi.transform.filters.fields.dewpoint",
    "anemoi.transform.filters.fields.sum",
    "anemoi.transform.filters.fields.orog_to_z",
)


def available() -> bool:
    return (REFERENCE_SRC / "anemoi" / "transform" / "spatial.py").exists()


def load() -> dict[str, object]:
    """→ {short name: module} of the reference's hot-path modules."""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    for p in (str(REPO), str(REPO / "anemoi-transform_b200"), str(STUBS), str(REFERENCE_SRC)):
        if p not in sys.path:
            sys.path.insert(0, p)
    mods = {}
    for name in HOT_PATH_MODULES:
        mods[name.rsplit(".", 1)[1]] = importlib.import_module(name)
    return mods
