"""Filters that convert groups of matching fields (u with v, q with t, …).

Reference `filters/fields/matching.py`: `MatchingSpec` 35-81, `MatchingFieldsFilter` 90-311.
Output ordering is the reference's (`_transform` 212-246): fields that take no part come
first, in input order; then, per group in first-seen order, the returned inputs followed by
the yielded outputs.

Addition: a subclass may implement `forward_batch(groups)` / `backward_batch(groups)`
returning, per group, the list of output fields — all groups then go through one device
pass instead of one numpy call per group.
"""

from __future__ import annotations

import logging
from abc import abstractmethod
from dataclasses import dataclass, replace
from inspect import signature
from itertools import chain
from typing import Any, Callable, Iterable, Iterator, Literal

import numpy as np

from .fields import new_field_from_numpy, new_fieldlist_from_list
from .filter import Filter
from .grouping import GroupByParam

LOG = logging.getLogger(__name__)


def _as_names(x: str | Iterable[str]) -> tuple[str, ...]:
    if isinstance(x, str):
        return (x,)
    try:
        return tuple(x)
    except TypeError as e:
        raise TypeError(f"Expected str or iterable, got {type(x)}") from e


@dataclass(frozen=True, slots=True)
class MatchingSpec:
    select: Literal["param"] = "param"
    forward: tuple[str, ...] = ()
    backward: tuple[str, ...] = ()
    return_inputs: Literal["all", "none"] | tuple[str, ...] = "none"
    vertical: bool = False

    def __post_init__(self) -> None:
        if self.select != "param":
            raise NotImplementedError("Only 'select=param' is supported for now.")
        object.__setattr__(self, "forward", _as_names(self.forward))
        object.__setattr__(self, "backward", _as_names(self.backward))
        if self.return_inputs not in ("all", "none"):
            object.__setattr__(self, "return_inputs", _as_names(self.return_inputs))
            all_params = set(self.forward) | set(self.backward)
            if not set(self.return_inputs).issubset(all_params):
                raise ValueError(f"Returned input names must subset {all_params}")

    def update_return_inputs(self, return_inputs) -> "MatchingSpec":
        if return_inputs not in ("all", "none"):
            return_inputs = _as_names(return_inputs)
        if return_inputs == self.return_inputs:
            return self
        return replace(self, return_inputs=return_inputs)

    def inputs(self, direction: Literal["forward", "backward"]) -> tuple[str, ...]:
        if self.return_inputs == "all":
            return tuple(getattr(self, direction))
        if self.return_inputs == "none":
            return ()
        return self.return_inputs


class MatchingFieldsFilter(Filter):
    MATCHING: MatchingSpec

    def __init_subclass__(cls, **kwargs: Any) -> None:
        super().__init_subclass__(**kwargs)
        if not isinstance(getattr(cls, "MATCHING", None), MatchingSpec):
            raise TypeError(f"Class {cls.__name__} must define a 'MATCHING' attribute of type MatchingSpec.")

        def check(method: Callable, expected: set[str]) -> None:
            missing = expected - set(signature(method).parameters)
            if missing:
                raise ValueError(f"{method}: missing parameters {missing}")

        fwd, bwd = set(cls.MATCHING.forward), set(cls.MATCHING.backward)
        check(cls.__init__, fwd | bwd)
        check(cls.forward_transform, fwd)
        check(cls.backward_transform, bwd)

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if hasattr(self, "return_inputs"):
            self.MATCHING = self.MATCHING.update_return_inputs(self.return_inputs)
        for direction in ("forward", "backward"):
            params = getattr(self.MATCHING, direction)
            inputs = self.MATCHING.inputs(direction=direction)
            if inputs and params and not set(inputs).issubset(params):
                LOG.warning(
                    f"Some {direction} inputs will not be returned because they are not in the filter parameters: "
                    f"{set(inputs) - set(params)}"
                )

    # ------------------------------------------------------------------ public API ----
    def forward(self, data: Any) -> Any:
        return self._run("forward", data)

    def backward(self, data: Any) -> Any:
        return self._run("backward", data)

    @abstractmethod
    def forward_transform(self, *fields: Any) -> Iterator[Any]: ...

    def backward_transform(self, *fields: Any) -> Iterator[Any]:
        raise NotImplementedError("Backward transformation not implemented.")

    def new_field_from_numpy(self, array: np.ndarray, *, template: Any, **kwargs: Any) -> Any:
        return new_field_from_numpy(array, template=template, **kwargs)

    def new_fieldlist_from_list(self, fields: list[Any]) -> Any:
        return new_fieldlist_from_list(fields)

    # ------------------------------------------------------------------ machinery -----
    def _run(self, direction: str, data: Any) -> Any:
        names = getattr(self.MATCHING, direction)                 # argument names, e.g. u_component
        params = [getattr(self, name) for name in names]          # metadata values, e.g. "u"
        returned = self.MATCHING.inputs(direction=direction)

        present = set(data.metadata(self.MATCHING.select))
        if not set(params).issubset(present):
            LOG.warning(
                "Please ensure your filter is configured to match the input variables metadata "
                f"current mismatch between inputs {present} and filter metadata {params}"
            )

        result: list[Any] = []
        groups = [dict(zip(names, g, strict=True)) for g in GroupByParam(params).iterate(data, other=result.append)]

        batch = getattr(self, f"{direction}_batch", None)
        if batch is not None and groups:
            produced = batch(groups)
        else:
            transform = getattr(self, f"{direction}_transform")
            produced = [list(transform(**g)) for g in groups]

        for g, outputs in zip(groups, produced, strict=True):
            result.extend(chain((g[name] for name in returned if name in g), outputs))
        return self.new_fieldlist_from_list(result)
