"""`rescale` / `convert` — reference `filters/fields/rescale.py:19-111`.

`x * scale + offset` forward, `(x - offset) / scale` backward, on every field of the selected
param in one device pass (kernels: AT_EPI_AFFINE / AT_EPI_AFFINE_INV in csrc/epilogue.cuh —
multiply then add, never fused, so float32 and float64 results are numpy's bit for bit).
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any

from ... import _cabi
from ...filter import SingleFieldFilter
from . import filter_registry
from .pointwise import NO_COL, device_field, run_epilogue


class Rescaler:
    """Host-side statement of the two formulas (reference rescale.py:19-29)."""

    def __init__(self, scale: float, offset: float):
        self.scale = scale
        self.offset = offset

    def forward(self, x):
        return x * self.scale + self.offset

    def backward(self, x):
        return (x - self.offset) / self.scale


class RescaleMixin(ABC):
    param: str
    rescaler: Rescaler
    forward_units = None
    backward_units = None

    @abstractmethod
    def prepare_filter(self):
        raise NotImplementedError("prepare_filter must be implemented by subclasses.")

    def forward_select(self):
        return {"param": self.param}

    def forward_transform(self, param: Any) -> Any:
        """Apply the forward transformation (x to ax+b)."""
        return self.forward_transform_batch([param])[0]

    def backward_transform(self, param: Any) -> Any:
        """Apply the backward transformation (ax+b to x)."""
        return self.backward_transform_batch([param])[0]

    def forward_transform_batch(self, fields: list[Any]) -> list[Any]:
        out = run_epilogue(_cabi.EPI_AFFINE, fields, [NO_COL] * len(fields), pa=float(self.rescaler.scale), pb=float(self.rescaler.offset))
        return [device_field(out, i, f, param=self.param, units=self.forward_units) for i, f in enumerate(fields)]

    def backward_transform_batch(self, fields: list[Any]) -> list[Any]:
        out = run_epilogue(_cabi.EPI_AFFINE_INV, fields, [NO_COL] * len(fields), pa=float(self.rescaler.scale), pb=float(self.rescaler.offset))
        return [device_field(out, i, f, param=self.param) for i, f in enumerate(fields)]


class Rescale(RescaleMixin, SingleFieldFilter):
    """A filter to rescale a parameter from a scale and an offset, and back."""

    required_inputs = ("scale", "offset", "param")

    def prepare_filter(self):
        self.rescaler = Rescaler(self.scale, self.offset)


class Convert(RescaleMixin, SingleFieldFilter):
    """A filter to convert a parameter in a given unit to another unit, and back (uses pint
    to derive the scale and offset, like the reference: rescale.py:92-106)."""

    required_inputs = ("unit_in", "unit_out", "param")

    def prepare_filter(self):
        import pint

        ureg = pint.UnitRegistry()
        self.forward_units = self.unit_out
        self.backward_units = self.unit_in
        x1, x2 = 0.0, 1.0
        y1 = ureg.Quantity(x1, self.unit_in).to(self.unit_out).magnitude
        y2 = ureg.Quantity(x2, self.unit_in).to(self.unit_out).magnitude
        scale = (y2 - y1) / (x2 - x1)
        offset = y1 - scale * x1
        self.rescaler = Rescaler(scale, offset)


filter_registry.register("rescale", Rescale)
filter_registry.register("convert", Convert)
