"""End-to-end host pipeline (at_pipeline_regrid: host fields in → host fields out)."""

import numpy as np
import pytest
from conftest import assert_same_values
from scipy.sparse import csr_array

from anemoi_transform_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_fields,chunk", [(1, 4), (7, 4), (64, 16), (130, 32), (33, 128)])
def test_host_pipeline_bit_exact(cuda, n_fields, chunk):
    from anemoi_transform_b200.device import CsrMatrix, HostPipeline

    t_lat, t_lon = syn.octahedral(48)
    d, i, p, shape = syn.bilinear_matrix(2.0, t_lat, t_lon)
    csr = CsrMatrix(d, i, p, shape)
    m = csr_array((d, i, p), shape=shape)
    fields = [syn.synthetic_field("t", shape[1], s, 0.001 if s % 5 == 0 else 0.0) for s in range(n_fields)]
    pipe = HostPipeline(csr, chunk_fields=chunk)
    out = pipe.regrid(fields)
    assert len(out) == n_fields and out[0].dtype == np.float32 and out[0].shape == (shape[0],)
    assert_same_values(np.stack(out), np.stack([m @ f for f in fields]), "host pipeline")
    # pinned host buffers (what bench.py uses) and re-use of the same pipeline object
    pinned_in = cuda.from_numpy(np.stack(fields)).pin_memory()
    pinned_out = cuda.empty((n_fields, shape[0]), dtype=cuda.float32).pin_memory()
    pipe.regrid(list(pinned_in.numpy()), list(pinned_out.numpy()))
    assert_same_values(pinned_out.numpy(), np.stack(out), "pinned buffers")
    pipe.close()


def test_no_device_memory_leak_over_repeated_calls(cuda, tmp_path):
    """Handles (matrix, epilogue programs, kNN indices, pipelines) are released: free HBM after 25
    rounds of filters and spatial calls equals free HBM after the first three."""
    import gc

    from anemoi_transform_b200 import ekd, spatial
    from anemoi_transform_b200.filters import create_filter_by_name as F
    from anemoi_transform_b200.source import FieldListSource

    s_lat, s_lon = syn.regular_latlon(2.0)
    t_lat, t_lon = syn.octahedral(24)
    matrix = str(tmp_path / "m.npz")
    syn.save_regrid_npz(matrix, *syn.bilinear_matrix(2.0, t_lat, t_lon), s_lat, s_lon, t_lat, t_lon)
    fields = [dict(param=p, levelist=850, values=syn.synthetic_field(p, s_lat.size, k), latitudes=s_lat, longitudes=s_lon) for k, p in enumerate(("u", "v", "q", "t"))]
    src = FieldListSource(dataset=ekd.from_source("list-of-dicts", fields))
    lam, glob = syn.rotated_lam(30, 30, 0.2, 50.0, 10.0), syn.octahedral(32)

    def once():
        pipe = src | F("regrid", matrix=matrix) | F("uv_to_ddff") | F("q_to_r") | F("clip", param="r", minimum=0.0, maximum=100.0)
        out = [f.to_numpy() for f in pipe]
        out += [f.to_numpy() for f in F("rescale", param="t", scale=2.0, offset=1.0).forward(src.forward(None))]
        spatial.cutout_mask(*lam, *glob)
        spatial.global_on_lam_mask(*lam, *glob)
        spatial.nearest_grid_points(*glob, *lam)
        return len(out)

    def free_mib():
        gc.collect()
        cuda.cuda.synchronize()
        cuda.cuda.empty_cache()
        return cuda.cuda.mem_get_info()[0] / 2**20

    for _ in range(3):
        once()
    before = free_mib()
    for _ in range(25):
        once()
    assert before - free_mib() < 32
