"""Matching fields that differ only by `param` — the contract of the reference's
`grouping/__init__.py:55-137`, implemented independently.

A field's identity is its "mars" namespace — or, when a field has none, every metadata key
except latitudes / longitudes / values — with `param` (and `variable`) taken out.  Fields
whose param is not wanted go to `other`, in input order; the rest are collected per identity,
identities in first-seen order, and every identity must end up with one field per wanted param
(`ValueError("Missing component…")`, grouping/__init__.py:135; a second field with the same
identity and param is a `ValueError("Duplicate component…")`).
"""

from __future__ import annotations

import logging
from typing import Any, Callable, Iterable, Iterator

LOG = logging.getLogger(__name__)

_NOT_IDENTITY = ("latitudes", "longitudes", "values")


def _lost(field: Any) -> None:
    raise ValueError(f"Lost field {field}")


def _flatten(params: Iterable[Any]) -> list[str]:
    """['u', ['v', ('w',)]] → ['u', 'v', 'w'] (iterative, depth first, order kept)."""
    out: list[str] = []
    pending = [iter(params)]
    while pending:
        for item in pending[-1]:
            if isinstance(item, (list, tuple)):
                pending.append(iter(item))
                break
            out.append(item)
        else:
            pending.pop()
    return out


def grouping_key(field: Any, extract: list[str], remove: list[str] | None = None) -> tuple[dict[str, Any], dict[str, Any]]:
    """→ (identity of the field without the extracted / removed keys, {extracted key: value})."""
    identity = dict(field.metadata(namespace="mars") or {})
    if not identity:
        names = [k for k in field.metadata().keys() if k not in _NOT_IDENTITY]
        if not names:
            raise NotImplementedError(f"GroupByParam: {field} has no sufficient metadata")
        identity = {k: field.metadata(k) for k in names}
    taken = {k: identity.pop(k) if k in identity else field.metadata().get(k, default=None) for k in extract}
    for k in remove or ():
        identity.pop(k, None)
    return identity, taken


class GroupByParam:
    """`iterate(fields, other=…)` yields one tuple of fields per identity, ordered as `params`."""

    def __init__(self, params: Any) -> None:
        self.params = _flatten(params if isinstance(params, (list, tuple)) else [params])

    def _get_groups(self, data: Iterable[Any], *, other: Callable[[Any], None] = _lost) -> None:
        assert callable(other), type(other)
        wanted = set(self.params)
        self.groups: dict[frozenset, dict[str, Any]] = {}
        self.groups_params: set[str] = set()
        for field in data:
            identity, taken = grouping_key(field, ["param"], ["variable"])
            param = taken["param"]
            if param not in wanted:
                other(field)
                continue
            members = self.groups.setdefault(frozenset(identity.items()), {})
            if param in members:
                raise ValueError(f"Duplicate component {param} for {frozenset(identity.items())}")
            members[param] = field
            self.groups_params.add(param)
        LOG.info(f"Params groups: {self.groups_params}")

    def iterate(self, data: Iterable[Any], *, other: Callable[[Any], None] = _lost) -> Iterator[tuple[Any, ...]]:
        self._get_groups(data, other=other)
        for members in self.groups.values():
            if len(members) != len(self.params):
                raise ValueError(f"Missing component. Want {sorted(self.params)}, got {sorted(members.keys())}")
            yield tuple(members[p] for p in self.params)
