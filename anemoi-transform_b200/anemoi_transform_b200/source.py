"""Source base class (reference `source.py`) — a Transform that ignores its input."""

from __future__ import annotations

from typing import Any

from .registry import Registry
from .transform import Transform

source_registry = Registry(__name__)


class Source(Transform):
    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"


@source_registry.register("fieldlist")
class FieldListSource(Source):
    """Wraps an in-memory FieldList so it can head a pipeline: `FieldListSource(fl) | regrid`."""

    def __init__(self, *, dataset: Any) -> None:
        assert dataset is not None, "Dataset cannot be None"
        self.ds = dataset

    def forward(self, *args: Any, **kwargs: Any) -> Any:
        return self.ds
