from anemoi_transform_b200.ekd import ArrayField  # noqa: F401
