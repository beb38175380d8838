from anemoi_transform_b200.filters import create_filter, create_filter_by_name, filter_registry  # noqa: F401
