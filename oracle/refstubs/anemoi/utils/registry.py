"""anemoi.utils.registry.Registry for the imported reference: the same small registry the
product uses (explicit imports register the factories; no directory scanning)."""
from anemoi_transform_b200.registry import Registry  # noqa: F401
