"""Field wrappers — new data / new metadata / new coordinates on top of a template field.

Behaviour follows the reference `fields.py`: `NewDataField` 139-205, `NewLatLonField`
318-372, `NewMetadataField` 441-535, `GeoMetadata` 208-315, the factories 645-738 and
`FieldSelection` 767-797.  Outputs of the regrid filter are wrapped exactly as
`NewLatLonField(NewMetadataField(NewDataField(template, array), **md), lat, lon)`
(regrid.py:312) so downstream consumers see the same metadata / geography.

`DeviceColumnField` is the one addition: a field whose values live in a column of a
`DeviceBatch` in HBM; `to_numpy()` downloads on demand, and the filters of this package
recognise it to keep a pipeline device-resident between filters.
"""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from .ekd import Geography, SimpleFieldList

LOG = logging.getLogger(__name__)

MISSING_METADATA = object()

_FORWARDED_QUIETLY = ("mars_area", "mars_grid", "to_numpy", "metadata", "shape", "grid_points", "handle")


def new_fieldlist_from_list(fields: list[Any]) -> SimpleFieldList:
    return SimpleFieldList(fields)


def new_empty_fieldlist() -> SimpleFieldList:
    return SimpleFieldList([])


class WrappedField:
    """Forwards everything it does not override to the wrapped field."""

    def __init__(self, field: Any) -> None:
        self._field = field

    def __getattr__(self, name: str) -> Any:
        if name in ("clone", "copy"):
            raise AttributeError(f"{self}: forwarding of `{name}` is not supported")
        if name.startswith("__") or name == "_field":
            raise AttributeError(name)
        if name not in _FORWARDED_QUIETLY:
            LOG.warning(f"{self}: forwarding `{name}`")
        return getattr(self._field, name)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self._field!r}, {self._repr_specific()})"

    def _repr_specific(self) -> str:
        return f"(No specific representation for {self.__class__.__name__})"

    def clone(self, **kwargs: Any) -> "NewClonedField":
        return NewClonedField(self, **kwargs)

    def __iter__(self) -> Any:
        raise NotImplementedError(f"{self}: iterating is not supported")


class NewDataField(WrappedField):
    """Template metadata, new values (a host numpy array owned by the field)."""

    def __init__(self, field: Any, data: np.ndarray) -> None:
        super().__init__(field)
        self._data = data
        self.shape = data.shape

    @property
    def values(self) -> np.ndarray:
        return self.to_numpy(flatten=True)

    def to_numpy(self, flatten: bool = False, dtype: type | None = None, index: Any | None = None) -> np.ndarray:
        data = self._data
        if dtype is not None:
            data = data.astype(dtype)
        if flatten:
            data = data.flatten()
        if index is not None:
            data = data[index]
        return data

    def _repr_specific(self) -> str:
        return f"(shape={self._data.shape})"


class DeviceColumnField(WrappedField):
    """Template metadata, values = column `col` of a point-major `DeviceBatch` in HBM."""

    def __init__(self, field: Any, batch: Any, col: int, shape: tuple[int, ...] | None = None) -> None:
        super().__init__(field)
        self._batch = batch
        self._col = int(col)
        self.shape = tuple(shape) if shape is not None else (batch.n_points,)

    @property
    def batch(self) -> Any:
        return self._batch

    @property
    def column(self) -> int:
        return self._col

    @property
    def values(self) -> np.ndarray:
        return self.to_numpy(flatten=True)

    def to_numpy(self, flatten: bool = False, dtype: type | None = None, index: Any | None = None) -> np.ndarray:
        # each call hands out a fresh array (reference NewDataField.to_numpy(flatten=True) copies)
        data = self._batch.take_column(self._col)
        if dtype is not None:
            data = data.astype(dtype)
        if not flatten:
            data = data.reshape(self.shape)
        if index is not None:
            data = data[index]
        return data

    def _repr_specific(self) -> str:
        return f"(device column {self._col} of {self._batch.n_points} points)"


class GeoMetadata(Geography):
    """Geography of a field whose coordinates were replaced."""

    def __init__(self, owner: Any) -> None:
        self.owner = owner

    def shape(self) -> tuple[int, ...]:
        return (len(self.owner._latitudes),)

    def resolution(self) -> str:
        return "unknown"

    def mars_area(self) -> list[float]:
        lat, lon = self.owner._latitudes, self.owner._longitudes
        return [np.amax(lat), np.amin(lon), np.amin(lat), np.amax(lon)]

    def mars_grid(self) -> None:
        return None

    def latitudes(self, dtype: type | None = None) -> np.ndarray:
        return self.owner._latitudes if dtype is None else self.owner._latitudes.astype(dtype)

    def longitudes(self, dtype: type | None = None) -> np.ndarray:
        return self.owner._longitudes if dtype is None else self.owner._longitudes.astype(dtype)

    def x(self, dtype: type | None = None) -> None:
        raise NotImplementedError()

    def y(self, dtype: type | None = None) -> None:
        raise NotImplementedError()

    def _unique_grid_id(self) -> None:
        raise NotImplementedError()

    def projection(self) -> None:
        return None

    def bounding_box(self) -> None:
        raise NotImplementedError()

    def gridspec(self) -> None:
        raise NotImplementedError()


class NewLatLonField(WrappedField):
    """Template values and metadata, new point coordinates."""

    def __init__(self, field: Any, latitudes: np.ndarray, longitudes: np.ndarray) -> None:
        super().__init__(field)
        self._latitudes = latitudes
        self._longitudes = longitudes

    def grid_points(self) -> tuple[np.ndarray, np.ndarray]:
        return self._latitudes, self._longitudes

    def to_latlon(self, flatten: bool = True) -> dict[str, np.ndarray]:
        assert flatten
        return dict(lat=self._latitudes, lon=self._longitudes)

    def metadata(self, *args: Any, **kwargs: Any) -> Any:
        metadata = self._field.metadata(*args, **kwargs)
        if hasattr(metadata, "geography"):
            metadata.geography = GeoMetadata(self)
        return metadata


class _MetadataView:
    """What `field.metadata()` returns for a field with overridden keys."""

    def __init__(self, owner: "NewMetadataField") -> None:
        self._owner = owner
        inner = owner._field.metadata()
        self.geography = getattr(inner, "geography", None)

    def get(self, key: str, default: Any = None) -> Any:
        value = self._owner.mapping(key, self._owner._field)
        if value is not MISSING_METADATA:
            return value
        return self._owner._field.metadata().get(key, default)

    def keys(self):
        return self._owner._field.metadata().keys()

    def __getitem__(self, key: str) -> Any:
        value = self._owner.mapping(key, self._owner._field)
        if value is not MISSING_METADATA:
            return value
        return self._owner._field.metadata()[key]

    def override(self, *args: Any, **kwargs: Any) -> Any:
        return self._owner._field.metadata().override(*args, **kwargs)


class NewMetadataField(WrappedField):
    """Template values, selected metadata keys overridden."""

    def __init__(self, field: Any, **kwargs: Any) -> None:
        super().__init__(field)
        self.kwargs = kwargs

    def mapping(self, key: str, field: Any) -> Any:
        return self.kwargs.get(key, MISSING_METADATA)

    def metadata(self, *args: Any, **kwargs: Any) -> Any:
        if not args and not kwargs:
            return _MetadataView(self)
        if kwargs.get("namespace"):
            assert len(args) == 0, (args, kwargs)
            ns = dict(self._field.metadata(**kwargs))
            for k in list(ns.keys()):
                m = self.mapping(k, self._field)
                if m is not MISSING_METADATA:
                    ns[k] = m
            return ns

        def one(key: str) -> Any:
            value = self.mapping(key, self._field)
            if value is MISSING_METADATA:
                return self._field.metadata(key, **kwargs)
            if callable(value):
                return value(self, key, self._field.metadata())
            return value

        result = [one(a) for a in args]
        return result[0] if len(result) == 1 else tuple(result)

    def _repr_specific(self) -> str:
        return f"(metadata={self.kwargs})"


class NewClonedField(WrappedField):
    def __init__(self, field: Any, **metadata: Any) -> None:
        super().__init__(field)
        self._metadata = metadata

    def metadata(self, *args: Any, **kwargs: Any) -> Any:
        if len(args) == 1 and args[0] in self._metadata:
            value = self._metadata[args[0]]
            if callable(value):
                value = self._metadata[args[0]] = value(self._field, args[0], self._field.metadata())
            return value
        return self._field.metadata(*args, **kwargs)

    def _repr_specific(self) -> str:
        return f"(metadata={self._metadata})"


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

    Walks through metadata / coordinate wrappers (they do not change values) and stops at
    the first wrapper that owns data.
    """
    f = field
    while True:
        if isinstance(f, DeviceColumnField):
            # an offloaded batch (DeviceBatch.offload) is a host field again
            return (f.batch, f.column) if getattr(f.batch, "resident", True) else None
        if isinstance(f, NewDataField):
            return None
        if isinstance(f, WrappedField):
            f = f._field
            continue
        return None


class FieldSelection:
    """Which fields a single-field filter applies to (keys: param, levelist)."""

    ALLOWED_KEYS = {"param", "levelist"}

    def __init__(self, **kwargs: Any):
        self._spec = kwargs
        if not set(self._spec).issubset(self.ALLOWED_KEYS):
            raise ValueError(f"Invalid keys in spec: {tuple(self._spec)} - only {self.ALLOWED_KEYS} are allowed.")
        for key, value in list(self._spec.items()):
            if isinstance(value, (str, int, float, bool)):
                self._spec[key] = (value,)
            elif value is None or (isinstance(value, (list, tuple)) and len(value) == 0):
                del self._spec[key]
            elif not isinstance(value, (list, tuple)):
                raise ValueError(f"Invalid value for key {key}: {value}")
        self._all = len(self._spec) == 0

    def match(self, field: Any) -> bool:
        if self._all:
            return True
        try:
            return all(field.metadata(key) in values for key, values in self._spec.items())
        except KeyError:
            return False
