from anemoi_transform_b200.source import source_registry  # noqa: F401
