"""Host-side mirror of the reference interface: registry names, constructor validation,
grouping / ordering, reversed transforms, pipelines.  No GPU: compute calls must fail loudly."""

import numpy as np
import pytest

from anemoi_transform_b200 import _cabi, ekd
from anemoi_transform_b200.fields import FieldSelection, new_field_from_latitudes_longitudes, new_field_from_numpy
from anemoi_transform_b200.filter import DispatchingFilter, SingleFieldFilter
from anemoi_transform_b200.filters import create_filter, create_filter_by_name, filter_registry
from anemoi_transform_b200.grouping import GroupByParam
from anemoi_transform_b200.matching import MatchingFieldsFilter, MatchingSpec
from anemoi_transform_b200.registry import Registry
from anemoi_transform_b200.source import FieldListSource
from anemoi_transform_b200.transform import ReversedTransform, Transform
from anemoi_transform_b200.workflows import Pipeline

MD = {"latitudes": [10.0, 0.0, -10.0], "longitudes": [20, 40.0], "valid_datetime": "2018-08-01T09:00:00Z"}


def fl(*specs):
    return ekd.from_source("list-of-dicts", [dict(param=p, levelist=lev, values=np.full((3, 2), float(v)), **MD) for p, lev, v in specs])


def test_registry_names_and_aliases():
    # reference: regrid.py:87, uv_to_ddff.py:130-131, q_to_r.py:84-85, clipper.py:18, clip.py:35, apply_mask.py:39, mask.py:35
    for name in ("regrid", "uv_to_ddff", "ddff_to_uv", "q_to_r", "r_to_q", "clip_fields", "clip", "apply_mask_fields", "mask"):
        assert name in filter_registry.registered
    assert filter_registry.lookup("clipper") is filter_registry.lookup("clip")
    assert filter_registry.lookup("apply_mask") is filter_registry.lookup("mask")
    assert filter_registry.lookup("nope", return_none=True) is None
    with pytest.raises(ValueError):
        filter_registry.lookup("nope")


def test_registry_api():
    r = Registry("pkg")

    @r.register("a", aliases=["b"])
    class A:
        def __init__(self, x=1):
            self.x = x

    assert r.create("a", x=2).x == 2 and r.create("b").x == 1
    assert r.from_config("a").x == 1 and r.from_config({"a": {"x": 5}}).x == 5 and r.from_config({"_type": "a", "x": 7}).x == 7
    assert r.registered == ["a"] and r.aliases() == {"a": ["b"]} and r.factories == {"a": A} and r.package == "pkg"
    with pytest.raises(AssertionError):
        r.register("a", A)
    with pytest.raises(ValueError):
        r.from_config({"a": {}, "c": {}})


def test_create_filter_sets_context():
    f = create_filter_by_name("uv_to_ddff", context="ctx")
    assert f.context == "ctx"
    f = create_filter("ctx2", {"clip": {"param": "t", "minimum": 0}})
    assert f.context == "ctx2"


def test_constructor_validation_messages():
    with pytest.raises(ValueError, match="At least one value for minimum or maximum"):
        create_filter_by_name("clip_fields", param="t")
    with pytest.raises(TypeError, match="Missing required input"):
        create_filter_by_name("clip_fields", minimum=0)
    with pytest.raises(ValueError, match="Unknown input"):
        create_filter_by_name("clip_fields", param="t", minimum=0, bogus=1)
    with pytest.raises(ValueError, match="Exactly one of `path` or `mask_param`"):
        create_filter_by_name("apply_mask_fields", mask_value=0)
    with pytest.raises(ValueError, match="Exactly one of `path` or `mask_param`"):
        create_filter_by_name("apply_mask", path="some_file", mask_param="lsm", mask_value=0)
    with pytest.raises(ValueError, match="Exactly one of `mask_value` or `threshold`"):
        create_filter_by_name("apply_mask", mask_param="lsm")
    with pytest.raises(ValueError, match="Invalid threshold operator"):
        create_filter_by_name("apply_mask", mask_param="lsm", threshold=1, threshold_operator="~")
    with pytest.raises(AssertionError, match="Radians"):
        create_filter_by_name("uv_to_ddff", radians=True)
    with pytest.raises(NotImplementedError, match="only 'nearest'"):
        from anemoi_transform_b200.filters.fields.regrid import ScipyKDTreeNearestNeighbours

        ScipyKDTreeNearestNeighbours(in_grid=None, out_grid=None, method="linear")
    grid = {"latitudes": np.zeros(2), "longitudes": np.zeros(2)}
    with pytest.raises(ValueError, match="out_grid is required"):
        ScipyKDTreeNearestNeighbours(in_grid=grid, out_grid=None, method="nearest")
    with pytest.raises(TypeError):  # None kwargs are dropped before construction (regrid.py:503-516)
        create_filter_by_name("regrid", method="nearest", in_grid=grid, out_grid=None)


def test_interpolator_precedence():
    from anemoi_transform_b200.filters.fields.regrid import _interpolator

    assert _interpolator(matrix="m", mask="k", method="nearest") == "MIRMatrix"
    assert _interpolator(mask="k", method="nearest") == "MaskedRegrid"
    assert _interpolator(method="nearest") == "ScipyKDTreeNearestNeighbours"
    assert _interpolator(method="linear") == "EarthkitRegrid"


def test_reversed_and_pipeline_and_not_reversible():
    class Plus(Transform):
        def __init__(self, n=1):
            self.n = n

        def forward(self, x):
            return x + self.n

        def backward(self, x):
            return x - self.n

    class OneWay(Transform):
        def forward(self, x):
            return x

    assert Plus.reversed(n=3).forward(10) == 7 and isinstance(Plus(2).reverse(), ReversedTransform)
    p = Plus(1) | Plus(2)
    assert isinstance(p, Pipeline) and p.forward(0) == 3 and p.backward(3) == 0 and p(0) == 3
    with pytest.raises(NotImplementedError, match="is not reversible"):
        OneWay().backward(1)
    assert Plus(1).patch_data_request({"a": 1}) == {"a": 1}


def test_field_wrappers_and_selection():
    f = fl(("t", 850, 1))[0]
    g = new_field_from_numpy(np.arange(6.0), template=f, param="tt")
    assert g.metadata("param") == "tt" and g.metadata("levelist") == 850 and g.metadata().get("param") == "tt"
    assert g.metadata("param", "levelist") == ("tt", 850)
    assert g.to_numpy(flatten=True) is not g.to_numpy(flatten=True)  # fresh copies
    h = new_field_from_latitudes_longitudes(g, np.array([1.0, 2.0]), np.array([3.0, 4.0]))
    assert h.grid_points()[0].tolist() == [1.0, 2.0] and h.metadata("param") == "tt"
    assert h.metadata().geography.latitudes().tolist() == [1.0, 2.0]
    assert FieldSelection(param="t").match(f) and not FieldSelection(param="q").match(f)
    assert FieldSelection(param=["q", "t"], levelist=850).match(f) and FieldSelection().match(f)
    assert not FieldSelection(levelist=500).match(fl(("t", 850, 1))[0])
    with pytest.raises(ValueError, match="Invalid keys"):
        FieldSelection(step=1)


def test_grouping_order_and_errors():
    data = fl(("t", 850, 0), ("u", 850, 1), ("z", 500, 2), ("v", 850, 3), ("u", 500, 4), ("v", 500, 5))
    other = []
    groups = list(GroupByParam(["u", "v"]).iterate(data, other=other.append))
    assert [f.metadata("param") for f in other] == ["t", "z"]
    assert [(u.metadata("levelist"), v.metadata("levelist")) for u, v in groups] == [(850, 850), (500, 500)]
    with pytest.raises(ValueError, match="Missing component"):
        list(GroupByParam(["u", "v"]).iterate(fl(("u", 850, 1)), other=other.append))
    with pytest.raises(ValueError, match="Duplicate component"):
        list(GroupByParam(["u", "v"]).iterate(fl(("u", 850, 1), ("u", 850, 2)), other=other.append))
    with pytest.raises(ValueError, match="Lost field"):
        list(GroupByParam(["u"]).iterate(fl(("t", 850, 1))))


def test_matching_filter_output_ordering_cpu_subclass():
    """A user-defined MatchingFieldsFilter (numpy only) keeps the reference's ordering:
    others first, then per group [returned inputs…, outputs…]  (matching.py:170-174, 242-246)."""

    class Summer(MatchingFieldsFilter):
        MATCHING = MatchingSpec(select="param", forward=("a", "b"), backward=("s",), return_inputs=("a",))

        def __init__(self, *, a="u", b="v", s="sum"):
            self.a, self.b, self.s = a, b, s
            super().__init__()

        def forward_transform(self, a, b):
            yield self.new_field_from_numpy(a.to_numpy() + b.to_numpy(), template=a, param=self.s)

        def backward_transform(self, s):
            yield s

    data = fl(("t", 850, 0), ("u", 850, 1), ("z", 500, 2), ("v", 850, 3), ("u", 500, 4), ("v", 500, 5))
    out = Summer().forward(data)
    assert [(f.metadata("param"), f.metadata("levelist")) for f in out] == [("t", 850), ("z", 500), ("u", 850), ("sum", 850), ("u", 500), ("sum", 500)]
    assert out[3].to_numpy().tolist() == np.full((3, 2), 4.0).tolist()
    with pytest.raises(ValueError, match="missing parameters"):

        class Bad(MatchingFieldsFilter):
            MATCHING = MatchingSpec(forward=("a",))

            def __init__(self):
                pass

            def forward_transform(self, x):
                yield x

    with pytest.raises(TypeError, match="must define a 'MATCHING'"):

        class Bad2(MatchingFieldsFilter):
            def forward_transform(self):
                yield None

    with pytest.raises(ValueError, match="Returned input names must subset"):
        MatchingSpec(forward=("a",), return_inputs=("zzz",))


def test_single_field_filter_and_dispatching_cpu_subclass():
    class Double(SingleFieldFilter):
        required_inputs = ("param",)

        def forward_select(self):
            return {"param": self.param}

        def forward_transform(self, field):
            return self.new_field_from_numpy(field.to_numpy() * 2, template=field)

    out = Double(param="t").forward(fl(("t", 850, 1), ("q", 850, 1)))
    assert out[0].to_numpy()[0, 0] == 2.0 and out[1].to_numpy()[0, 0] == 1.0
    with pytest.raises(NotImplementedError, match="backward transform not implemented"):
        Double(param="t").backward(fl(("t", 850, 1)))
    with pytest.raises(TypeError, match="must override at least one"):

        class Nothing(DispatchingFilter):
            pass

    clip = create_filter_by_name("clip", param="t", minimum=0)
    with pytest.raises(TypeError, match="No forward method"):
        clip.forward(42)
    with pytest.raises(NotImplementedError):
        create_filter_by_name("clip", minimum=0)  # tabular configuration


def test_compute_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for CPU-only boxes")
    data = fl(("u", 850, 1), ("v", 850, 2))
    with pytest.raises(_cabi.NativeLibraryError):
        (FieldListSource(dataset=data) | create_filter_by_name("uv_to_ddff")).forward(None)
    from anemoi_transform_b200 import spatial

    with pytest.raises(_cabi.NativeLibraryError):
        spatial.nearest_grid_points(np.zeros(3), np.zeros(3), np.zeros(2), np.zeros(2))


def test_product_never_imports_the_oracle():
    from conftest import REPO

    for path in (REPO / "anemoi-transform_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
