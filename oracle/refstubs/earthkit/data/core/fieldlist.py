from anemoi_transform_b200.ekd import Field, FieldList  # noqa: F401
