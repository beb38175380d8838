"""Group fields that differ only by `param` — reference `grouping/__init__.py:55-137`.

Key of a field = its "mars" namespace (or, when that is empty, every metadata key except
latitudes / longitudes / values) minus `param` (and `variable`).  Fields whose param is not
wanted go to `other` in input order; groups come out in first-seen order and must be
complete (`ValueError("Missing component…")`, grouping/__init__.py:135).
"""

from __future__ import annotations

import logging
from collections import defaultdict
from typing import Any, Callable, Iterator

LOG = logging.getLogger(__name__)


def _lost(f: Any) -> None:
    raise ValueError(f"Lost field {f}")


def _flatten(params) -> list[str]:
    flat: list[str] = []
    for p in params:
        if isinstance(p, (list, tuple)):
            flat.extend(_flatten(p))
        else:
            flat.append(p)
    return flat


def grouping_key(field: Any, extract: list[str], remove: list[str] | None = None):
    key = field.metadata(namespace="mars")
    key = dict(key) if key else {}
    if not key:
        meta_keys = [k for k in field.metadata().keys() if k not in ("latitudes", "longitudes", "values")]
        if not meta_keys:
            raise NotImplementedError(f"GroupByParam: {field} has no sufficient metadata")
        key = {k: field.metadata(k) for k in meta_keys}
    extracted = {}
    for k in extract:
        extracted[k] = key.pop(k, field.metadata().get(k, default=None))
    for k in remove or []:
        key.pop(k, None)
    return key, extracted


class GroupByParam:
    def __init__(self, params) -> None:
        if not isinstance(params, (list, tuple)):
            params = [params]
        self.params = _flatten(params)

    def _get_groups(self, data, *, other: Callable[[Any], None] = _lost) -> None:
        assert callable(other), type(other)
        self.groups: dict[frozenset, dict[str, Any]] = defaultdict(dict)
        self.groups_params = set()
        for f in data:
            key, extras = grouping_key(f, ["param"], ["variable"])
            param = extras["param"]
            if param not in self.params:
                other(f)
                continue
            key = frozenset(key.items())
            if param in self.groups[key]:
                raise ValueError(f"Duplicate component {param} for {key}")
            self.groups[key][param] = f
            self.groups_params.add(param)
        LOG.info(f"Params groups: {self.groups_params}")

    def iterate(self, data, *, other: Callable[[Any], None] = _lost) -> Iterator[tuple[Any, ...]]:
        self._get_groups(data, other=other)
        for group in self.groups.values():
            if len(group) != len(self.params):
                raise ValueError(f"Missing component. Want {sorted(self.params)}, got {sorted(group.keys())}")
            yield tuple(group[p] for p in self.params)
