"""Registry of the tabular (DataFrame) filters (reference `filters/tabular/__init__.py:11-13`).

The tabular / observation path is outside the hot path; the one filter provided here,
`assign_to_grid`, is the SURVEY §8(f) "next" row that reuses the device kNN."""

from ...registry import Registry

filter_registry = Registry(__name__)
