"""The C-ABI library loads on a CPU-only box and exports every symbol include/at_b200.h declares."""

import ctypes
import re

from conftest import REPO


def declared_symbols():
    text = (REPO / "include" / "at_b200.h").read_text()
    return sorted(set(re.findall(r"AT_API\s+(?:const\s+)?\w+\*?\s+(at_\w+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    names = declared_symbols()
    assert len(names) >= 28
    for must in ("at_csr_create", "at_spmm", "at_spmm_fused", "at_pointwise", "at_knn_query", "at_ball_mark", "at_cutout_classify", "at_pipeline_regrid"):
        assert must in names


def test_library_exports_every_declared_symbol(native_library):
    for name in declared_symbols():
        assert hasattr(native_library, name), f"libat_b200.so does not export {name}"


def test_python_binding_covers_the_header(native_library):
    from anemoi_transform_b200 import _cabi

    assert sorted(_cabi.PROTOTYPES) == declared_symbols()


def test_no_compute_without_gpu_is_reported_not_faked(native_library):
    """Host-only entry points work; anything needing a device reports an error code."""
    assert native_library.at_version() >= 100
    n = ctypes.c_int(-1)
    rc = native_library.at_device_count(ctypes.byref(n))
    import torch

    if not torch.cuda.is_available():
        assert rc != 0 and native_library.at_last_error()
    else:
        assert rc == 0 and n.value >= 1


def test_every_entry_point_cites_the_reference():
    text = (REPO / "include" / "at_b200.h").read_text()
    for ref in ("regrid.py:309-310", "regrid.py:281-285", "spatial.py:236-275", "spatial.py:189-233", "spatial.py:533-534", "spatial.py:93-97", "apply_mask.py:160-163", "regrid.py:380"):
        assert ref in text, ref
