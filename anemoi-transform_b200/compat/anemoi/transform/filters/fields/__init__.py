from anemoi_transform_b200.filters.fields import filter_registry  # noqa: F401
