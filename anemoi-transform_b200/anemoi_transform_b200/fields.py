"""Field overlays: new values, new metadata or new coordinates laid over a template field.

What the reference's wrappers do (`fields.py`: `NewDataField` 139-205, `GeoMetadata` 208-315,
`NewLatLonField` 318-372, `NewMetadataField` 441-535, `NewClonedField` 538-600, factories
645-738, `FieldSelection` 767-797) is the contract; the implementation here is this package's
own.  Outputs of the regrid filter are wrapped exactly as
`NewLatLonField(NewMetadataField(NewDataField(template, array), **md), lat, lon)`
(regrid.py:312), so a downstream consumer sees the same metadata and geography.

Every overlay is an `_Overlay`: it keeps the field it wraps in `_field` and answers only what
it changes; anything else is looked up on the wrapped field (`__getattr__`), with a warning for
names outside the field protocol — the reference logs the same, which is how accidental
pass-through is noticed.

`DeviceColumnField` is the addition: values that live in a column of a point-major
`DeviceBatch` in HBM.  `to_numpy()` brings them to the host on demand; the filters of this
package recognise such fields (`device_column_of`) and keep a pipeline resident between
filters.
"""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from .ekd import Geography, SimpleFieldList

LOG = logging.getLogger(__name__)

MISSING_METADATA = object()

#: the field protocol: forwarded to the wrapped field without a word
_PROTOCOL = frozenset({"to_numpy", "metadata", "shape", "grid_points", "mars_area", "mars_grid", "handle"})
#: never forwarded: a copy of the wrapped field would silently drop the overlay
_NOT_FORWARDED = frozenset({"clone", "copy"})


def new_fieldlist_from_list(fields: list[Any]) -> SimpleFieldList:
    return SimpleFieldList(fields)


def new_empty_fieldlist() -> SimpleFieldList:
    return SimpleFieldList([])


def _as_requested(values: np.ndarray, flatten: bool, dtype: Any, index: Any) -> np.ndarray:
    """The `to_numpy(flatten=, dtype=, index=)` options, applied in the reference's order
    (cast, then flatten — `flatten()` always copies —, then index)."""
    out = values if dtype is None else values.astype(dtype)
    if flatten:
        out = out.flatten()
    return out if index is None else out[index]


class WrappedField:
    """Base of the overlays: holds the wrapped field, forwards what it does not define."""

    def __init__(self, field: Any) -> None:
        self._field = field

    def __getattr__(self, name: str) -> Any:
        # reached only for names the overlay itself does not have
        if name == "_field" or name.startswith("__"):
            raise AttributeError(name)
        if name in _NOT_FORWARDED:
            raise AttributeError(f"{self}: forwarding of `{name}` is not supported")
        if name not in _PROTOCOL:
            LOG.warning(f"{self}: forwarding `{name}`")
        return getattr(self._field, name)

    def _repr_specific(self) -> str:
        return f"(No specific representation for {type(self).__name__})"

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self._field!r}, {self._repr_specific()})"

    def __iter__(self) -> Any:
        raise NotImplementedError(f"{self}: iterating is not supported")

    def clone(self, **metadata: Any) -> "NewClonedField":
        return NewClonedField(self, **metadata)


_Overlay = WrappedField


class NewDataField(_Overlay):
    """The template's metadata over a host array the field owns."""

    def __init__(self, field: Any, data: np.ndarray) -> None:
        _Overlay.__init__(self, field)
        self._data, self.shape = data, data.shape

    def to_numpy(self, flatten: bool = False, dtype: Any = None, index: Any = None) -> np.ndarray:
        return _as_requested(self._data, flatten, dtype, index)

    @property
    def values(self) -> np.ndarray:
        return self.to_numpy(flatten=True)

    def _repr_specific(self) -> str:
        return f"(shape={self.shape})"


class DeviceColumnField(_Overlay):
    """The template's metadata over column `col` of a point-major `DeviceBatch` in HBM."""

    def __init__(self, field: Any, batch: Any, col: int, shape: tuple[int, ...] | None = None) -> None:
        _Overlay.__init__(self, field)
        self._batch, self._col = batch, int(col)
        self.shape = (batch.n_points,) if shape is None else tuple(shape)

    batch = property(lambda self: self._batch)
    column = property(lambda self: self._col)

    def to_numpy(self, flatten: bool = False, dtype: Any = None, index: Any = None) -> np.ndarray:
        # the batch hands out a fresh host array per call — the caller may write into it, as it
        # may into the copy the reference's flatten() makes (apply_mask.py:184-185)
        column = self._batch.take_column(self._col)
        if dtype is not None:
            column = column.astype(dtype)
        if not flatten:
            column = column.reshape(self.shape)
        return column if index is None else column[index]

    @property
    def values(self) -> np.ndarray:
        return self.to_numpy(flatten=True)

    def _repr_specific(self) -> str:
        return f"(device column {self._col} of {self._batch.n_points} points)"


class GeoMetadata(Geography):
    """Geography of a field whose point coordinates were replaced: latitudes, longitudes, shape
    and the MARS area are known; a grid description is not (the points are arbitrary)."""

    def __init__(self, owner: Any) -> None:
        self.owner = owner

    def _coordinates(self, which: str, dtype: Any) -> np.ndarray:
        values = getattr(self.owner, which)
        return values if dtype is None else values.astype(dtype)

    def latitudes(self, dtype: Any = None) -> np.ndarray:
        return self._coordinates("_latitudes", dtype)

    def longitudes(self, dtype: Any = None) -> np.ndarray:
        return self._coordinates("_longitudes", dtype)

    def shape(self) -> tuple[int, ...]:
        return (len(self.owner._latitudes),)

    def mars_area(self) -> list[float]:
        lat, lon = self.owner._latitudes, self.owner._longitudes
        north, south, west, east = np.amax(lat), np.amin(lat), np.amin(lon), np.amax(lon)
        return [north, west, south, east]

    def resolution(self) -> str:
        return "unknown"

    def mars_grid(self) -> None:
        return None

    def projection(self) -> None:
        return None

    def _undefined(self, *args: Any, **kwargs: Any) -> None:
        raise NotImplementedError()

    x = y = bounding_box = gridspec = _unique_grid_id = _undefined


class NewLatLonField(_Overlay):
    """The template's values and metadata at new point coordinates."""

    def __init__(self, field: Any, latitudes: np.ndarray, longitudes: np.ndarray) -> None:
        _Overlay.__init__(self, field)
        self._latitudes, self._longitudes = latitudes, longitudes

    def grid_points(self) -> tuple[np.ndarray, np.ndarray]:
        return self._latitudes, self._longitudes

    def to_latlon(self, flatten: bool = True) -> dict[str, np.ndarray]:
        assert flatten
        return {"lat": self._latitudes, "lon": self._longitudes}

    def metadata(self, *args: Any, **kwargs: Any) -> Any:
        answer = self._field.metadata(*args, **kwargs)
        if hasattr(answer, "geography"):  # a whole metadata object: its geography is ours now
            answer.geography = GeoMetadata(self)
        return answer


class _MetadataView:
    """`field.metadata()` of a field with overridden keys: reads go through the overrides,
    everything else (key listing, `override`) is the wrapped field's."""

    def __init__(self, owner: "NewMetadataField") -> None:
        self._owner = owner
        self._inner = owner._field.metadata()
        self.geography = getattr(self._inner, "geography", None)

    def _overridden(self, key: str) -> Any:
        return self._owner.mapping(key, self._owner._field)

    def __getitem__(self, key: str) -> Any:
        value = self._overridden(key)
        return self._owner._field.metadata()[key] if value is MISSING_METADATA else value

    def get(self, key: str, default: Any = None) -> Any:
        value = self._overridden(key)
        return self._owner._field.metadata().get(key, default) if value is MISSING_METADATA else value

    def keys(self):
        return self._owner._field.metadata().keys()

    def override(self, *args: Any, **kwargs: Any) -> Any:
        return self._owner._field.metadata().override(*args, **kwargs)


class NewMetadataField(_Overlay):
    """The template's values with some metadata keys overridden."""

    def __init__(self, field: Any, **kwargs: Any) -> None:
        _Overlay.__init__(self, field)
        self.kwargs = kwargs

    def mapping(self, key: str, field: Any) -> Any:
        """The override for `key`, or MISSING_METADATA (subclasses may compute it)."""
        return self.kwargs.get(key, MISSING_METADATA)

    def _one(self, key: str, **kwargs: Any) -> Any:
        value = self.mapping(key, self._field)
        if value is MISSING_METADATA:
            return self._field.metadata(key, **kwargs)
        return value(self, key, self._field.metadata()) if callable(value) else value

    def metadata(self, *args: Any, **kwargs: Any) -> Any:
        if not args and not kwargs:
            return _MetadataView(self)
        if kwargs.get("namespace"):
            assert not args, (args, kwargs)
            namespace = dict(self._field.metadata(**kwargs))
            for key in namespace:
                value = self.mapping(key, self._field)
                if value is not MISSING_METADATA:
                    namespace[key] = value
            return namespace
        values = tuple(self._one(key, **kwargs) for key in args)
        return values[0] if len(values) == 1 else values

    def _repr_specific(self) -> str:
        return f"(metadata={self.kwargs})"


class NewClonedField(_Overlay):
    """`field.clone(key=value, …)`: single-key metadata reads answer from the clone's keys
    (callables are evaluated once, on first use)."""

    def __init__(self, field: Any, **metadata: Any) -> None:
        _Overlay.__init__(self, field)
        self._metadata = metadata

    def metadata(self, *args: Any, **kwargs: Any) -> Any:
        if len(args) != 1 or args[0] not in self._metadata:
            return self._field.metadata(*args, **kwargs)
        key = args[0]
        if callable(self._metadata[key]):
            self._metadata[key] = self._metadata[key](self._field, key, self._field.metadata())
        return self._metadata[key]

    def _repr_specific(self) -> str:
        return f"(metadata={self._metadata})"


# ---- factories (reference fields.py:645-738) ---------------------------------------------
def new_field_from_numpy(array: np.ndarray, *, template: Any, **metadata: Any) -> NewMetadataField:
    return NewMetadataField(NewDataField(template, array), **metadata)


def new_field_from_device_column(batch: Any, col: int, *, template: Any, shape=None, **metadata: Any) -> NewMetadataField:
    return NewMetadataField(DeviceColumnField(template, batch, col, shape), **metadata)


def new_field_with_metadata(template: Any, **metadata: Any) -> NewMetadataField:
    return NewMetadataField(template, **metadata)


def new_field_from_latitudes_longitudes(template: Any, latitudes: np.ndarray, longitudes: np.ndarray) -> NewLatLonField:
    return NewLatLonField(template, latitudes, longitudes)


def device_column_of(field: Any):
    """→ (batch, column) when `field`'s VALUES are a device column, else None.

    Metadata and coordinate overlays do not change values, so the walk looks through them and
    stops at the first overlay that owns data."""
    node = field
    while isinstance(node, WrappedField):
        if isinstance(node, DeviceColumnField):
            # an offloaded batch (DeviceBatch.offload) is a host field again
            return (node.batch, node.column) if getattr(node.batch, "resident", True) else None
        if isinstance(node, NewDataField):
            return None
        node = node._field
    return None


class FieldSelection:
    """Which fields a single-field filter applies to: `param` and / or `levelist`, each a value
    or a list of values; no key at all selects every field."""

    ALLOWED_KEYS = {"param", "levelist"}

    def __init__(self, **kwargs: Any):
        unknown = set(kwargs) - self.ALLOWED_KEYS
        if unknown:
            raise ValueError(f"Invalid keys in spec: {tuple(kwargs)} - only {self.ALLOWED_KEYS} are allowed.")
        self._spec: dict[str, tuple] = {}
        for key, wanted in kwargs.items():
            if wanted is None or (isinstance(wanted, (list, tuple)) and not wanted):
                continue  # an absent or empty entry does not restrict
            if isinstance(wanted, (str, int, float, bool)):
                wanted = (wanted,)
            elif not isinstance(wanted, (list, tuple)):
                raise ValueError(f"Invalid value for key {key}: {wanted}")
            self._spec[key] = tuple(wanted)
        self._all = not self._spec

    def match(self, field: Any) -> bool:
        try:
            return all(field.metadata(key) in wanted for key, wanted in self._spec.items())
        except KeyError:
            return False
