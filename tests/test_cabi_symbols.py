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


def test_a_plain_c_program_can_consume_the_header_and_the_library(native_library, tmp_path):
    """The boundary is a C ABI: a C99 translation unit includes include/at_b200.h, links
    libat_b200.so and calls it (host-only entry points here; argument validation must answer
    with a status and a message, never crash)."""
    import shutil
    import subprocess

    from anemoi_transform_b200 import _cabi

    if shutil.which("gcc") is None:
        import pytest

        pytest.skip("gcc not available")
    lib = _cabi.library_path()
    src = tmp_path / "consumer.c"
    src.write_text(
        """
#include <stdio.h>
#include <string.h>
#include "at_b200.h"
int main(void) {
    at_csr_t* csr = NULL;
    int32_t indptr[2] = {0, 1};
    int32_t indices[1] = {7}; /* column 7 of a 1 x 3 matrix: invalid */
    float data[1] = {1.0f};
    int rc;
    if (at_version() < 100) return 1;
    rc = at_csr_create(1, 3, 1, indptr, AT_I32, indices, AT_I32, data, AT_F32, &csr);
    if (rc != AT_ERR_INVALID || csr != NULL) return 2;
    if (strstr(at_last_error(), "indices") == NULL) return 3;
    if (at_spmm(NULL, NULL, AT_F32, 0, NULL, AT_F32, 0, 0, 0, NULL) == AT_OK) return 4;
    if (at_csr_destroy(NULL) != AT_OK || at_knn_destroy(NULL) != AT_OK || at_pipeline_destroy(NULL) != AT_OK) return 5;
    printf("consumer ok: %s\\n", at_last_error());
    return 0;
}
"""
    )
    exe = tmp_path / "consumer"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", f"-I{REPO / 'include'}", str(src), "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"]
    build = subprocess.run(cmd, capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and "consumer ok" in run.stdout, (run.returncode, run.stdout, run.stderr)
