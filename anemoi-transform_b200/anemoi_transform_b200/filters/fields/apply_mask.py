"""`apply_mask_fields` — reference `filters/fields/apply_mask.py:39-245`.

Sets the selected fields to NaN where a mask holds.  The mask is `mask_values == mask_value`
or `OP(mask_values, threshold)` (apply_mask.py:160-163) of either a file (`path`) or a field
of the pipeline (`mask_param`, consumed unless `return_mask`).  The comparison runs on the
device (`at_compare_mask`), the masking in one `at_pointwise` pass (AT_COL_MASK).
"""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from ... import _cabi
from ...batching import fields_to_batch
from ...device import compare_mask, require_cuda
from ...fields import FieldSelection, new_fieldlist_from_list
from ...filter import Filter
from . import filter_registry
from .pointwise import device_field, run_epilogue

LOG = logging.getLogger(__name__)

# spelling -> at_compare_mask op code (0 ==, 1 !=, 2 >, 3 >=, 4 <, 5 <=)
OPERATORS = {
    ">": 2,
    "<": 4,
    "==": 0,
    "!=": 1,
    ">=": 3,
    "<=": 5,
    "gt": 2,
    "lt": 4,
    "eq": 0,
    "ne": 1,
    "ge": 3,
    "le": 5,
}


@filter_registry.register("apply_mask_fields")
class MaskVariable(Filter):
    """Mask variables using a mask from a file or from a field in the pipeline."""

    def __init__(
        self,
        *,
        path: str | None = None,
        mask_param: str | None = None,
        mask_value: float | None = None,
        threshold: float | None = None,
        threshold_operator: str = ">",
        rename: str | None = None,
        param: str | list[str] | None = None,
        return_mask: bool = False,
    ) -> None:
        self.path = path
        self.mask_param = mask_param
        self.mask_value = mask_value
        self.threshold = threshold
        self.threshold_operator = threshold_operator
        self.rename = rename
        self.param = param if not isinstance(param, str) else [param]
        self.return_mask = return_mask
        self.mask = None
        self.prepare_filter()
        self._forward_selection = FieldSelection(**self.forward_select())

    def prepare_filter(self) -> None:
        if (self.path is None) == (self.mask_param is None):
            raise ValueError("Exactly one of `path` or `mask_param` must be provided.")
        if (self.mask_value is None) == (self.threshold is None):
            raise ValueError("Exactly one of `mask_value` or `threshold` must be provided.")
        if self.threshold is not None and self.threshold_operator not in OPERATORS:
            raise ValueError(
                f"Invalid threshold operator: {self.threshold_operator}. "
                f"Valid operators are: {', '.join(OPERATORS.keys())}."
            )
        if self.path is not None:
            if self.path.endswith(".npy"):
                values = np.load(self.path)
            else:
                from ... import ekd

                values = ekd.from_source("file", self.path)[0].to_numpy(flatten=True)
            self.mask = self._compute_mask(values)

    def _compute_mask(self, mask_values: Any):
        """uint8 CUDA tensor: 1 where the field must become NaN."""
        torch = require_cuda()
        if isinstance(mask_values, np.ndarray):
            v = torch.from_numpy(np.ascontiguousarray(mask_values.reshape(-1))).cuda()
        else:
            v = mask_values
        if self.threshold is not None:
            return compare_mask(v, OPERATORS[self.threshold_operator], self.threshold)
        return compare_mask(v, 0, self.mask_value)

    def forward_select(self) -> dict[str, Any]:
        if self.param is not None:
            return {"param": self.param}
        return {}

    def _separate_mask_and_fields(self, fields: Any):
        if self.mask_param is None:
            return self.mask, list(fields)
        mask_field = None
        remaining = []
        for field in fields:
            if field.metadata("param") == self.mask_param:
                if mask_field is None:
                    mask_field = field  # first instance is the mask
                if not self.return_mask:
                    continue
            remaining.append(field)
        if mask_field is None:
            raise ValueError(f"Mask parameter '{self.mask_param}' not found in input data.")
        column = fields_to_batch([mask_field])
        return self._compute_mask(column.data[:, 0]), remaining

    def forward_transform(self, field: Any) -> Any:
        return self._mask_fields([field])[0]

    def _mask_fields(self, fields: list[Any]) -> list[Any]:
        n_points = int(np.prod(fields[0].shape))
        if self.mask is None or int(self.mask.shape[0]) != n_points:
            have = None if self.mask is None else int(self.mask.shape[0])
            raise IndexError(f"boolean index did not match indexed array: mask has {have} points, field has {n_points}")
        out = run_epilogue(_cabi.EPI_PLAIN, fields, [(0.0, 0.0, 0.0, _cabi.COL_MASK)] * len(fields), row_mask=self.mask)
        result = []
        for i, f in enumerate(fields):
            metadata = {}
            if self.rename is not None:
                metadata["param"] = f"{f.metadata('param')}_{self.rename}"
            # the reference flattens the values before masking (apply_mask.py:184)
            result.append(device_field_flat(out, i, f, **metadata))
        return result

    def forward(self, fields: Any) -> Any:
        self.mask, fields = self._separate_mask_and_fields(fields)
        picked = [i for i, f in enumerate(fields) if self._forward_selection.match(f)]
        if picked:
            for i, out in zip(picked, self._mask_fields([fields[i] for i in picked]), strict=True):
                fields[i] = out
        return new_fieldlist_from_list(fields)


def device_field_flat(batch, col, template, **metadata):
    from ...fields import new_field_from_device_column

    batch, col = batch.locate(col)
    return new_field_from_device_column(batch, col, template=template, shape=None, **metadata)
