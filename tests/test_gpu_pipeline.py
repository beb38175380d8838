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
