"""Top-level filter registry — reference `filters/__init__.py:13-64`.

The field-filter registry is merged into the top-level one with duplicate detection
(`_merge_registries`, reference 22-33); the dispatchers `clip` / `mask` register directly at
the top level.  Tabular (pandas) filters are outside the hot path and not provided; the
dispatchers raise a clear error if asked for a tabular configuration.
"""

from __future__ import annotations

from typing import Any

from ..registry import Registry
from .fields import filter_registry as fields_filter_registry
from .tabular import filter_registry as tabular_filter_registry

filter_registry = Registry(__name__, entry_point_group="anemoi.transform.filters")

# importing the modules runs their registration decorators
from .fields import apply_mask as _apply_mask  # noqa: E402,F401
from .fields import clipper as _clipper  # noqa: E402,F401
from .fields import cos_sin_from_rad as _cos_sin_from_rad  # noqa: E402,F401
from .fields import cos_sin_mean_wave_direction as _cos_sin_mwd  # noqa: E402,F401
from .fields import dewpoint as _dewpoint  # noqa: E402,F401
from .fields import icon_refinement_level as _icon_refinement_level  # noqa: E402,F401
from .fields import impute_nans as _impute_nans_fields  # noqa: E402,F401
from .fields import lnsp_to_sp as _lnsp_to_sp  # noqa: E402,F401
from .fields import orog_to_z as _orog_to_z  # noqa: E402,F401
from .fields import q_to_r as _q_to_r  # noqa: E402,F401
from .fields import regrid as _regrid  # noqa: E402,F401
from .fields import remove_nans as _remove_nans_fields  # noqa: E402,F401
from .fields import rescale as _rescale  # noqa: E402,F401
from .fields import sum as _sum  # noqa: E402,F401
from .fields import uv_to_ddff as _uv_to_ddff  # noqa: E402,F401
from .tabular import assign_to_grid as _assign_to_grid  # noqa: E402,F401
from .tabular import superob as _superob  # noqa: E402,F401


def _merge_registries() -> None:
    for source in (fields_filter_registry, tabular_filter_registry):
        for name, factory in source.factories.items():
            try:
                filter_registry.register(name, factory, aliases=source.aliases().get(name, None))
            except AssertionError as e:
                raise AssertionError(f"Duplicate filter name: {name} in {source.package} registry") from e


def create_filter_by_name(name: str, *, context: Any = None, **config: Any):
    """Create a filter from its registered name and keyword configuration."""
    f = filter_registry.create(name, **config)
    f.context = context
    return f


def create_filter(context: Any, config: Any):
    """Create a filter from a YAML-style configuration (`"name"` or `{"name": {…}}`)."""
    f = filter_registry.from_config(config)
    f.context = context
    return f


_merge_registries()

from . import clip as _clip  # noqa: E402,F401
from . import geopotential_to_height as _geopotential_to_height  # noqa: E402,F401
from . import impute_nans as _impute_nans  # noqa: E402,F401
from . import mask as _mask  # noqa: E402,F401
from . import remove_nans as _remove_nans  # noqa: E402,F401

__all__ = ["filter_registry", "create_filter", "create_filter_by_name"]
