"""`sum` — reference `filters/fields/sum.py:22-122`.

Sums a set of params per matching key (the mars namespace minus `param`, optionally minus
`levelist`), in field order, one device pass for all keys (`at_sum_cols`: `s = c0; s += c1;
…` in the batch's dtype).  The summed inputs are consumed, every other field passes through
first, the sums follow in first-seen key order (sum.py:88-118).
"""

from __future__ import annotations

import logging
from collections import defaultdict
from collections.abc import Hashable
from typing import Any

from ...batching import fields_to_batch
from ...device import DeviceBatch, results_are_host_bound, sum_cols
from ...fields import new_field_from_device_column, new_fieldlist_from_list
from ...filter import Filter
from . import filter_registry

LOG = logging.getLogger(__name__)


@filter_registry.register("sum")
class Sum(Filter):
    """Computes the sum over a set of variables."""

    def __init__(self, *, params: list[str], output: str, ignore_level: bool = False):
        self.params = params
        self.output = output
        self.ignore_level = ignore_level

    def forward(self, fields: Any) -> Any:
        result = []
        needed_fields: dict[tuple[Hashable, ...], dict[str, Any]] = defaultdict(dict)
        for f in fields:
            key = dict(f.metadata(namespace="mars"))
            param = key.pop("param", None)
            if self.ignore_level:
                ll = key.pop("levelist", None)
                LOG.debug(f"Removing levelist ({ll}) from matching key for variable: {param}")
            if param is None:
                param = f.metadata("param")
            if param in self.params:
                key = tuple(key.items())
                if param in needed_fields[key]:
                    raise ValueError(f"Duplicate field {param} for {key}")
                needed_fields[key][param] = f
            else:
                result.append(f)

        groups = []
        for keys, values in needed_fields.items():
            if len(values) != len(self.params):
                raise ValueError("Missing fields")
            groups.append(list(values.values()))
        if groups:
            n_terms = len(self.params)
            batch = fields_to_batch([f for g in groups for f in g])
            out = DeviceBatch(sum_cols(batch.data, list(range(len(groups) * n_terms)), len(groups), n_terms), len(groups))
            if results_are_host_bound():
                out.prefetch()
            for i, g in enumerate(groups):
                # the reference sums flattened arrays (sum.py:110)
                result.append(new_field_from_device_column(out, i, template=g[0], shape=None, param=self.output))
        return new_fieldlist_from_list(result)

    def backward(self, data: Any) -> Any:
        raise NotImplementedError("Sum filter is not reversible")
