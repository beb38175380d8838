"""Transform base classes — same public behaviour as the reference `transform.py`.

Mirrors (reference src/anemoi/transform/transform.py): `Transform.forward/backward/
reverse/__call__/__or__/patch_data_request` (47-172), the metaclass `.reversed` class
property that builds `ReversedTransform(cls(**kw))` (27-44) and `ReversedTransform` (175-244).
"""

from __future__ import annotations

from abc import ABC, ABCMeta, abstractmethod
from typing import Any, Callable


class _ReversibleMeta(ABCMeta):
    """Gives every Transform subclass a `.reversed(**kwargs)` constructor."""

    @property
    def reversed(cls) -> Callable[..., "ReversedTransform"]:
        def make_reversed(*args: Any, **kwargs: Any) -> "ReversedTransform":
            return ReversedTransform(cls(*args, **kwargs))

        make_reversed.__doc__ = cls.__doc__
        make_reversed.__name__ = f"reversed_{cls.__name__}"
        return make_reversed


class Transform(ABC, metaclass=_ReversibleMeta):
    """Abstract base of every filter, source and workflow."""

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"

    def __call__(self, data: Any) -> Any:
        return self.forward(data)

    @abstractmethod
    def forward(self, data: Any) -> Any: ...

    def backward(self, data: Any) -> Any:
        raise NotImplementedError(f"{self} is not reversible.")

    def reverse(self) -> "Transform":
        return ReversedTransform(self)

    def __or__(self, other: "Transform") -> "Transform":
        from .workflows import workflow_registry

        return workflow_registry.create("pipeline", filters=[self, other])

    def patch_data_request(self, data_request: dict) -> dict:
        return data_request

    def reversed(self, *args: Any, **kwargs: Any) -> "Transform":
        return self.__class__.reversed(*args, **kwargs)


class ReversedTransform(Transform):
    """Swaps forward and backward of the wrapped transform."""

    def __init__(self, filter: Transform) -> None:
        self.filter = filter

    def __repr__(self) -> str:
        return f"Reversed({self.filter})"

    def forward(self, x: Any) -> Any:
        return self.filter.backward(x)

    def backward(self, x: Any) -> Any:
        return self.filter.forward(x)

    def patch_data_request(self, data_request: dict) -> dict:
        return self.filter.patch_data_request(data_request)
