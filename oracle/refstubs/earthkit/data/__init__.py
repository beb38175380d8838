"""Stub of earthkit.data: the array-backed field model of anemoi_transform_b200.ekd."""
from anemoi_transform_b200.ekd import ArrayField, Field, FieldList, SimpleFieldList, from_source  # noqa: F401
from . import core, indexing  # noqa: F401
