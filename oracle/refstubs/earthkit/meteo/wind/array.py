from oracle.pointwise import polar_to_xy, xy_to_polar  # noqa: F401
