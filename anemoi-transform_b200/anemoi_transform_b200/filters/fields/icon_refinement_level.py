"""`icon_refinement_level` — reference `filters/fields/icon_refinement_level.py:25-85`.

Nearest-neighbour interpolation of every field to the cell centres of an ICON grid (optionally
only the cells up to a refinement level).  The reference computes the indices once with
`anemoi.utils.grids.nearest_grid_points` (xyz + cKDTree) and then fancy-indexes field by field
(`data[..., self.nearest_grid_points]`, :77); here the search is the device kNN and the gather
one `at_gather_rows` over the whole FieldList — the machinery of `regrid(method="nearest")`.
"""

from __future__ import annotations

from typing import Any

from ...fields import new_fieldlist_from_list
from ...filter import Filter
from ...grids import icon_grid
from . import filter_registry
from .regrid import ScipyKDTreeNearestNeighbours


@filter_registry.register("icon_refinement_level")
class IconRefinement(Filter):
    """Interpolate the input to an ICON grid."""

    def __init__(self, *, grid: str, refinement_level_c: int | None) -> None:
        self.grid = grid
        self.refinement_level_c = refinement_level_c
        self.latitudes, self.longitudes = icon_grid(self.grid, self.refinement_level_c)
        # all fields are assumed to share the first field's grid, as in the reference (:66)
        self._nearest = ScipyKDTreeNearestNeighbours(in_grid=None, out_grid=dict(latitudes=self.latitudes, longitudes=self.longitudes), method="nearest")

    @property
    def nearest_grid_points(self) -> Any:
        """Indices into the input grid, one per ICON cell (None before the first `forward`)."""
        idx = self._nearest.nearest_grid_points
        return None if idx is None else idx.cpu().numpy()

    def forward(self, fields: Any) -> Any:
        fields = list(fields)
        if not fields:
            return new_fieldlist_from_list([])
        return new_fieldlist_from_list(self._nearest.regrid_batch(fields))
