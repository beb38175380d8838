from anemoi_transform_b200.spatial import *  # noqa: F401,F403
from anemoi_transform_b200.spatial import (  # noqa: F401
    cropping_mask,
    cutout_mask,
    global_on_lam_mask,
    latlon_to_xyz,
    nearest_grid_points,
    outline,
    thinning_mask,
    xyz_to_latlon,
)
