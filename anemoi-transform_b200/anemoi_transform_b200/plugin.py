"""Plugging the B200 filters into anemoi-transform's own registry.

Two routes:

* **entry points** — `pyproject.toml` publishes the factories below in the
  `anemoi.transform.filters` group under `b200_*` names; the reference's registry
  (`anemoi.transform.filters.filter_registry`, filters/__init__.py:19, an
  `anemoi.utils.registry.Registry`) loads that group for names it does not know, so a recipe
  can say `b200_regrid: {matrix: …}` next to stock filters;
* **`install()`** — replaces the reference's built-in factories for the hot-path names
  (`regrid`, `uv_to_ddff`, `q_to_r`, `clip`, `mask`, …) with the B200 ones, so unchanged
  recipes and `create_filter(context, config)` calls run on the GPU.

The filters of this package only rely on the field protocol (`to_numpy`, `metadata`,
`grid_points`, iteration), take the reference's constructor arguments and raise its errors, so
they run inside the reference's `Pipeline` / `a | b` / `Transform.reversed` machinery as they
do inside this package's.
"""

from __future__ import annotations

import logging
from typing import Any, Callable

LOG = logging.getLogger(__name__)


def _factories() -> dict[str, Callable[..., Any]]:
    """Registered name (the reference's) → B200 factory."""
    from .filters import filter_registry

    names = (
        "regrid", "uv_to_ddff", "ddff_to_uv", "q_to_r", "r_to_q", "clip", "clipper", "clip_fields", "mask", "apply_mask",
        "apply_mask_fields", "rescale", "convert", "lnsp_to_sp", "sp_to_lnsp", "impute_nans", "remove_nans", "cos_sin_from_rad",
        "cos_sin_mean_wave_direction", "r_to_d", "d_to_r", "sum", "orog_to_z", "z_to_orog",
    )  # fmt: skip
    return {name: filter_registry.lookup(name) for name in names if filter_registry.lookup(name, return_none=True) is not None}


def __getattr__(name: str) -> Any:
    """`anemoi_transform_b200.plugin:<name>` — what the entry points of pyproject.toml load."""
    table = _factories()
    if name in table:
        return table[name]
    raise AttributeError(name)


def install(registry: Any = None, names: list[str] | None = None, prefix: str = "") -> list[str]:
    """Register the B200 factories in `registry` (default: the reference's
    `anemoi.transform.filters.filter_registry`) under `prefix + name`, replacing a factory
    already registered under that name.  → the names registered.

    `install()` makes stock recipes run on the B200 path; `install(prefix="b200_")` adds the
    filters next to the stock ones (what the entry points do without any call)."""
    if registry is None:
        from anemoi.transform.filters import filter_registry as registry  # the reference's

    done = []
    for name, factory in _factories().items():
        if names is not None and name not in names:
            continue
        target = prefix + name
        _replace(registry, target, factory)
        done.append(target)
    LOG.info("anemoi_transform_b200: registered %s", done)
    return done


def _replace(registry: Any, name: str, factory: Callable[..., Any]) -> None:
    """`registry.register(name, factory)`, dropping an existing registration of `name` first:
    registries refuse duplicates (anemoi.utils.registry asserts on them)."""
    for attribute in ("_factories", "factories", "registered"):  # anemoi-utils keeps a dict under one of these
        table = getattr(registry, attribute, None)
        if isinstance(table, dict) and name in table:
            del table[name]
            break
    alias_of = getattr(registry, "_alias_of", None)
    if isinstance(alias_of, dict) and name in alias_of:  # an alias of a stock filter: detach it
        del alias_of[name]
    registry.register(name, factory)
