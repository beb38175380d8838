#!/usr/bin/env python
"""A regrid pipeline on the B200 path, written exactly as it would be with the reference:

    python examples/regrid_pipeline.py            (needs a CUDA device and the built library)

1° lat-lon → O48 with a locally built bilinear matrix in the `make-regrid-file` npz schema,
then wind speed / direction, relative humidity, a clip and a land mask.  `regrid` and the four
pointwise filters run as ONE kernel launch (fusion.py); fields stay in HBM until `to_numpy()`.
"""

import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "anemoi-transform_b200"))

from anemoi_transform_b200 import ekd  # noqa: E402
from anemoi_transform_b200 import synthetic as syn  # noqa: E402
from anemoi_transform_b200.filters import create_filter_by_name as create_filter  # noqa: E402
from anemoi_transform_b200.source import FieldListSource  # noqa: E402


def main():
    s_lat, s_lon = syn.regular_latlon(1.0)
    t_lat, t_lon = syn.octahedral(48)
    with tempfile.TemporaryDirectory() as tmp:
        matrix = str(Path(tmp) / "1deg-to-o48.npz")
        syn.save_regrid_npz(matrix, *syn.bilinear_matrix(1.0, t_lat, t_lon), s_lat, s_lon, t_lat, t_lon)

        fields = []
        for level in (500, 850):
            for k, param in enumerate(("u", "v", "q", "t")):
                fields.append(dict(param=param, levelist=level, values=syn.synthetic_field(param, s_lat.size, 10 * level + k), latitudes=s_lat, longitudes=s_lon))
        fields.append(dict(param="lsm", levelist=0, values=(np.random.default_rng(0).uniform(size=s_lat.size) > 0.7).astype(np.float32), latitudes=s_lat, longitudes=s_lon))
        source = FieldListSource(dataset=ekd.from_source("list-of-dicts", fields))

        pipeline = (
            source
            | create_filter("regrid", matrix=matrix)
            | create_filter("uv_to_ddff")
            | create_filter("q_to_r")
            | create_filter("clip", param="r", minimum=0.0, maximum=100.0)
            | create_filter("apply_mask", mask_param="lsm", threshold=0.5, threshold_operator=">", param=["ws", "wdir"])
        )
        for field in pipeline:
            values = field.to_numpy(flatten=True)
            print(f"{field.metadata('param'):5s} level {field.metadata('levelist'):4d}  {values.shape[0]} points  mean {np.nanmean(values):10.4f}  NaNs {int(np.isnan(values).sum())}")


if __name__ == "__main__":
    main()
