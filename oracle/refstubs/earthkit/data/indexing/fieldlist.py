from anemoi_transform_b200.ekd import SimpleFieldList  # noqa: F401

FieldArray = SimpleFieldList
