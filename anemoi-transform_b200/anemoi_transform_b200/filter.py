"""Filter base classes — reference `filter.py`: `Filter` 29-32, `DispatchingFilter` 35-99,
`SingleFieldFilter` 102-202 (constructor validation messages included)."""

from __future__ import annotations

import logging
from abc import abstractmethod
from typing import Any, Callable

import numpy as np

from . import ekd
from .fields import FieldSelection, new_field_from_numpy, new_fieldlist_from_list
from .transform import Transform

try:  # tabular data is outside the hot path; dispatch on it only when pandas is there
    import pandas as pd

    _DataFrame = pd.DataFrame
except Exception:  # pragma: no cover

    class _DataFrame:  # type: ignore[no-redef]
        pass


LOG = logging.getLogger(__name__)


class Filter(Transform):
    """A transform that processes field data."""


class DispatchingFilter(Transform):
    """Routes FieldLists to `forward_fields` and DataFrames to `forward_tabular`."""

    def __init_subclass__(cls, **kwargs: Any) -> None:
        super().__init_subclass__(**kwargs)

        def overridden(name: str) -> bool:
            return getattr(cls, name) is not getattr(DispatchingFilter, name)

        if not (overridden("forward_fields") or overridden("forward_tabular")):
            raise TypeError(f"{cls.__name__} must override at least one of `forward_fields` or `forward_tabular`")
        for kind in ("fields", "tabular"):
            if overridden(f"backward_{kind}") and not overridden(f"forward_{kind}"):
                raise TypeError(f"{cls.__name__} overrides `backward_{kind}` but not `forward_{kind}`")

    def _route(self, direction: str, data: Any) -> Any:
        """Pick `<direction>_fields` for FieldLists, `<direction>_tabular` for DataFrames and the
        direction's fallback (which raises) for anything else."""
        if isinstance(data, ekd.FieldList):
            kind = "fields"
        elif isinstance(data, _DataFrame):
            kind = "tabular"
        else:
            kind = "fallback"
        return getattr(self, f"{direction}_{kind}")(data)

    def forward(self, data: Any) -> Any:
        return self._route("forward", data)

    def backward(self, data: Any) -> Any:
        return self._route("backward", data)

    # defaults: a kind a subclass does not implement ends in the direction's fallback
    def forward_fallback(self, data: Any) -> Any:
        raise TypeError(f"No forward method for {type(data)}")

    def backward_fallback(self, data: Any) -> Any:
        raise NotImplementedError(f"No backward method for {type(data)}")

    def forward_fields(self, data: Any) -> Any:
        return self.forward_fallback(data)

    def forward_tabular(self, data: Any) -> Any:
        return self.forward_fallback(data)

    def backward_fields(self, data: Any) -> Any:
        return self.backward_fallback(data)

    def backward_tabular(self, data: Any) -> Any:
        return self.backward_fallback(data)


class SingleFieldFilter(Filter):
    """Transforms fields one at a time; non-selected fields pass through unchanged.

    Subclasses declare `required_inputs` / `optional_inputs`; constructor kwargs become
    attributes.  Subclasses may also implement `forward_transform_batch(fields)` to process
    all selected fields in one device pass.
    """

    required_inputs: tuple[str, ...] | list[str] | None = None
    optional_inputs: dict[str, Any] = {}

    def __init__(self, **kwargs: Any) -> None:
        # configuration = declared defaults overlaid with what the caller gave
        self._config = {**self.optional_inputs, **kwargs}
        self._validate_inputs()
        self.prepare_filter()
        selections = {"_forward_selection": self.forward_select(), "_backward_selection": self.backward_select()}
        for attribute, spec in selections.items():
            setattr(self, attribute, FieldSelection(**spec))

    def prepare_filter(self) -> None:
        pass

    def forward_select(self) -> dict[str, Any]:
        return {}

    def backward_select(self) -> dict[str, Any]:
        return self.forward_select()

    @abstractmethod
    def forward_transform(self, field: Any) -> Any: ...

    def backward_transform(self, field: Any) -> Any:
        raise NotImplementedError("Field backward transform not implemented.")

    def new_field_from_numpy(self, array: np.ndarray, *, template: Any, **metadata: Any) -> Any:
        return new_field_from_numpy(array, template=template, **metadata)

    def _validate_inputs(self) -> None:
        """The constructor arguments against `required_inputs` / `optional_inputs` (the messages
        are the reference's, filter.py:165-178: its tests match on them)."""
        required = self.required_inputs
        if not required:
            return
        if not isinstance(required, (list, tuple)):
            raise TypeError("Required inputs must be a list or tuple.")
        given = set(self._config)
        missing = set(required) - given
        if missing:
            raise TypeError(f"Missing required input(s): '{missing}'.")
        unknown = given - set(required) - set(self.optional_inputs)
        if unknown:
            raise ValueError(f"Unknown input(s): '{unknown}'.")

    @property
    def config(self) -> dict[str, Any]:
        return self._config

    def __getattr__(self, name: str) -> Any:
        if name.startswith("__") or name == "_config":
            raise AttributeError(name)
        try:
            return self._config[name]
        except KeyError:
            raise AttributeError(name) from None

    def _apply(self, data: Any, selection: FieldSelection, one: Callable[[Any], Any], batch: Callable | None) -> Any:
        fields = list(data)
        picked = [i for i, f in enumerate(fields) if selection.match(f)]
        if batch is not None and picked:
            for i, out in zip(picked, batch([fields[i] for i in picked]), strict=True):
                fields[i] = out
        else:
            for i in picked:
                fields[i] = one(fields[i])
        return new_fieldlist_from_list(fields)

    def forward(self, data: Any) -> Any:
        return self._apply(data, self._forward_selection, self.forward_transform, getattr(self, "forward_transform_batch", None))

    def backward(self, data: Any) -> Any:
        return self._apply(data, self._backward_selection, self.backward_transform, getattr(self, "backward_transform_batch", None))
