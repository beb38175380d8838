"""`rescale` / `convert` — reference `filters/fields/rescale.py:19-111`.

Forward `x·scale + offset`, backward `(x − offset) / scale`, applied to every field of the
selected param in one device pass (AT_EPI_AFFINE / AT_EPI_AFFINE_INV in csrc/epilogue.cuh:
multiply then add, never fused, IEEE division — float32 and float64 results are numpy's bit
for bit).  `convert` derives scale and offset from two units with pint, like the reference.
"""

from __future__ import annotations

from typing import Any

from ... import _cabi
from ...filter import SingleFieldFilter
from . import filter_registry
from .pointwise import NO_COL, device_field, run_epilogue


class Rescaler:
    """The affine map and its inverse as plain Python (what the device kernels compute)."""

    def __init__(self, scale: float, offset: float):
        self.scale, self.offset = scale, offset

    def forward(self, x):
        return x * self.scale + self.offset

    def backward(self, x):
        return (x - self.offset) / self.scale


class _AffineFilter(SingleFieldFilter):
    """Shared device path of `Rescale` and `Convert`; subclasses set `self.rescaler` (and the
    unit names) in `prepare_filter`."""

    rescaler: Rescaler
    forward_units = None
    backward_units = None

    def forward_select(self) -> dict[str, Any]:
        return {"param": self.param}

    def _run(self, kind: int, fields: list[Any], **metadata: Any) -> list[Any]:
        pa, pb = float(self.rescaler.scale), float(self.rescaler.offset)
        out = run_epilogue(kind, fields, [NO_COL] * len(fields), pa=pa, pb=pb)
        return [device_field(out, i, f, param=self.param, **metadata) for i, f in enumerate(fields)]

    def forward_transform_batch(self, fields: list[Any]) -> list[Any]:
        return self._run(_cabi.EPI_AFFINE, fields, units=self.forward_units)

    def backward_transform_batch(self, fields: list[Any]) -> list[Any]:
        return self._run(_cabi.EPI_AFFINE_INV, fields)

    def forward_transform(self, param: Any) -> Any:
        return self.forward_transform_batch([param])[0]

    def backward_transform(self, param: Any) -> Any:
        return self.backward_transform_batch([param])[0]


# name kept for code that tests `isinstance(f, RescaleMixin)` (fusion planner)
RescaleMixin = _AffineFilter


class Rescale(_AffineFilter):
    """Rescale a parameter with a scale and an offset, and back."""

    required_inputs = ("scale", "offset", "param")

    def prepare_filter(self) -> None:
        self.rescaler = Rescaler(self.scale, self.offset)


class Convert(_AffineFilter):
    """Convert a parameter from `unit_in` to `unit_out`, and back (needs pint)."""

    required_inputs = ("unit_in", "unit_out", "param")

    def prepare_filter(self) -> None:
        import pint

        quantity = pint.UnitRegistry().Quantity
        at_zero = quantity(0.0, self.unit_in).to(self.unit_out).magnitude
        at_one = quantity(1.0, self.unit_in).to(self.unit_out).magnitude
        self.forward_units, self.backward_units = self.unit_out, self.unit_in
        self.rescaler = Rescaler(at_one - at_zero, at_zero)


filter_registry.register("rescale", Rescale)
filter_registry.register("convert", Convert)
