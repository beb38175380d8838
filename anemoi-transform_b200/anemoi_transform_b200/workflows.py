"""Workflow / Pipeline — reference `workflow.py:17-43`, `workflows/pipeline.py:18-65`."""

from __future__ import annotations

from typing import Any, Iterator

from .registry import Registry
from .transform import Transform

workflow_registry = Registry(__name__)


class Workflow(Transform):
    def __iter__(self) -> Iterator[Any]:
        return iter(self(None))

    def __call__(self, data: Any) -> Any:
        return self.forward(data)


@workflow_registry.register("pipeline")
class Pipeline(Workflow):
    """`a | b | c`: forward applies the filters in order, backward in reverse order."""

    def __init__(self, *, filters: list[Any]) -> None:
        self.filters = filters
        self._plan: list[Any] | None = None
        self._plan_key: tuple[int, ...] = ()

    def execution_plan(self) -> list[Any]:
        """The filters as they will run forward: nested pipelines flattened and
        `regrid | pointwise…` runs fused into one launch (see fusion.py).  Results are
        those of running `self.filters` one after the other."""
        from .fusion import flatten, fuse

        key = tuple(id(f) for f in flatten(self.filters))
        if self._plan is None or key != self._plan_key:
            self._plan, self._plan_key = fuse(self.filters), key
        return self._plan

    def forward(self, data: Any) -> Any:
        from .device import results_stay_on_device

        plan = self.execution_plan()
        for f in plan[:-1]:
            with results_stay_on_device():  # the next filter consumes them in HBM
                data = f.forward(data)
        for f in plan[-1:]:
            data = f.forward(data)
        return data

    def backward(self, data: Any) -> Any:
        for f in reversed(self.filters):
            data = f.backward(data)
        return data

    def __repr__(self) -> str:
        return f"Pipeline({self.filters})"
