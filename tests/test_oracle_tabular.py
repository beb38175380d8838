"""oracle/tabular.py against the imported reference (build container only)."""

import numpy as np
import pytest

from oracle import reference_import
from oracle import tabular as ot

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not reference_import.available(), reason="reference sources not present")]


def test_superob_restatement_equals_the_reference(monkeypatch):
    reference_import.load()
    import importlib

    ref_support = importlib.import_module("anemoi.transform.filters.tabular.support.superob")
    ref_filter = importlib.import_module("anemoi.transform.filters.tabular.superob")
    from anemoi_transform_b200 import synthetic as syn

    lat, lon = syn.octahedral(6)
    grid = np.column_stack([lat, np.where(lon > 180, lon - 360, lon)])
    monkeypatch.setattr(ref_filter, "define_grid", lambda name: grid)  # the named-grid download is not the arithmetic
    obs = ot.synthetic_observations(5000, seed=3)
    obs.loc[::501, "latitude"] = np.nan
    binned = ref_support.assign_nearest_grid(obs.dropna(subset=["latitude"]), grid, 3600)
    mine = ot.assign_nearest_grid(obs.dropna(subset=["latitude"]), grid, 3600)
    assert binned.equals(mine)
    want = ref_filter.SuperOb(grid="o6", timeslot_length=3600, columns_to_take_nearest=["date"], columns_to_groupby=["reporttype"]).forward(obs.copy())
    got = ot.superob(obs.copy(), grid, 3600, take_nearest=["date"], groupby=["reporttype"])
    assert list(want.columns) == list(got.columns) and len(want) == len(got)
    assert want.reset_index(drop=True).equals(got.reset_index(drop=True))
