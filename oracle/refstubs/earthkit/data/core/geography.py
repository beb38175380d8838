from anemoi_transform_b200.ekd import Geography  # noqa: F401
