"""Registry of the field filters (reference `filters/fields/__init__.py:11-13`)."""

from ...registry import Registry

filter_registry = Registry(__name__)
