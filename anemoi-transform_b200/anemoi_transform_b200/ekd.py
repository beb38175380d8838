"""The slice of the earthkit-data field model the hot path consumes.

The reference builds on `earthkit.data` (Field, FieldList, SimpleFieldList, ArrayField,
`from_source("list-of-dicts", …)`).  earthkit-data is not installed in the build / GPU
images, so this module provides array-backed stand-ins with the protocol listed in
SURVEY.md §8(b): `to_numpy(flatten, dtype, index)`, `metadata(*keys, namespace=, default=)`,
`metadata()` with `.get / .keys / [] / .geography / .override`, `grid_points()`, `shape`,
`values`; FieldList iteration, `len`, `[i]`, `.metadata(key)`, `.sel`, `.to_numpy()`.

When the real earthkit-data is importable the filters accept its fields unchanged — they
only rely on the protocol above (duck typing), exactly as the reference does
(e.g. filters/fields/regrid.py:309, matching.py:239, grouping/__init__.py:70-91).
"""

from __future__ import annotations

from typing import Any, Iterable, Iterator

import numpy as np

_MISSING = object()

# keys that describe geography / payload rather than identify a field
_NON_IDENTITY_KEYS = ("latitudes", "longitudes", "values")


class Geography:
    """Base class of geography objects (earthkit.data.core.geography.Geography)."""


class ArrayGeography(Geography):
    def __init__(self, latitudes: np.ndarray, longitudes: np.ndarray, shape: tuple[int, ...]):
        self._lat, self._lon, self._shape = latitudes, longitudes, shape

    def latitudes(self, dtype=None):
        return self._lat if dtype is None else self._lat.astype(dtype)

    def longitudes(self, dtype=None):
        return self._lon if dtype is None else self._lon.astype(dtype)

    def shape(self):
        return self._shape

    def resolution(self):
        return "unknown"

    def mars_grid(self):
        return None

    def mars_area(self):
        return [np.amax(self._lat), np.amin(self._lon), np.amin(self._lat), np.amax(self._lon)]


class Metadata:
    """Dict-backed metadata with an optional fake "mars" namespace."""

    MARS_KEYS = frozenset({"param", "levelist", "levtype", "type", "step", "date", "time", "number", "expver", "class", "stream", "domain"})

    def __init__(self, data: dict[str, Any], geography: Geography | None = None, mars: bool = False):
        self._data = dict(data)
        self.geography = geography
        self._mars = mars

    def get(self, key: str, default: Any = None) -> Any:
        return self._data.get(key, default)

    def keys(self):
        return self._data.keys()

    def items(self):
        return self._data.items()

    def __getitem__(self, key: str) -> Any:
        return self._data[key]

    def __contains__(self, key: str) -> bool:
        return key in self._data

    def as_namespace(self, namespace: str | None = None) -> dict[str, Any]:
        if self._mars and namespace == "mars":
            return {k: v for k, v in self._data.items() if k in self.MARS_KEYS}
        return {}

    def override(self, *args: Any, **kwargs: Any) -> "Metadata":
        d = dict(self._data)
        for a in args:
            d.update(a)
        d.update(kwargs)
        return Metadata(d, self.geography, self._mars)


class _ForeignMetadata(Metadata):
    """Adapter around a dict-like metadata object that implements `as_namespace` itself."""

    def __init__(self, inner: Any):
        super().__init__({k: inner[k] for k in inner.keys()}, getattr(inner, "geography", None), False)
        self._inner = inner

    def as_namespace(self, namespace: str | None = None) -> dict[str, Any]:
        return dict(self._inner.as_namespace(namespace))


class Field:
    """Base class of fields (earthkit.data.Field)."""


class ArrayField(Field):
    """A field backed by a numpy array and a metadata dict."""

    def __init__(self, array: Any, metadata: dict[str, Any] | Metadata, *, latitudes=None, longitudes=None, mars: bool = False):
        self._array = np.asarray(array)
        if isinstance(metadata, Metadata):
            self._metadata = metadata
        elif hasattr(metadata, "as_namespace"):
            # a dict-like metadata object that brings its own namespaces (earthkit's RawMetadata /
            # UserMetadata subclasses): keep its as_namespace
            self._metadata = _ForeignMetadata(metadata)
        else:
            geography = None
            if latitudes is not None and longitudes is not None:
                geography = ArrayGeography(np.asarray(latitudes), np.asarray(longitudes), self._array.shape)
            self._metadata = Metadata(metadata, geography, mars)

    @property
    def shape(self) -> tuple[int, ...]:
        return self._array.shape

    @property
    def values(self) -> np.ndarray:
        return self.to_numpy(flatten=True)

    def to_numpy(self, flatten: bool = False, dtype=None, index=None) -> np.ndarray:
        data = self._array
        if dtype is not None:
            data = data.astype(dtype)
        if flatten:
            data = data.flatten()
        if index is not None:
            data = data[index]
        return data

    def metadata(self, *keys: str, namespace: str | None = None, default: Any = _MISSING, **kwargs: Any) -> Any:
        if namespace is not None:
            assert not keys, (keys, namespace)
            return self._metadata.as_namespace(namespace)
        if not keys:
            return self._metadata
        out = []
        for k in keys:
            if k in self._metadata:
                out.append(self._metadata[k])
            elif default is not _MISSING:
                out.append(default)
            else:
                raise KeyError(k)
        return out[0] if len(out) == 1 else tuple(out)

    def grid_points(self) -> tuple[np.ndarray, np.ndarray]:
        g = self._metadata.geography
        if g is None:
            raise ValueError("field has no geography")
        return g.latitudes(), g.longitudes()

    def to_latlon(self, flatten: bool = True) -> dict[str, np.ndarray]:
        lat, lon = self.grid_points()
        return dict(lat=lat, lon=lon)

    def __repr__(self) -> str:
        ident = {k: v for k, v in self._metadata.items() if k not in _NON_IDENTITY_KEYS}
        return f"ArrayField({ident}, shape={self.shape})"


class FieldList:
    """Base class of field lists (earthkit.data.FieldList)."""


class SimpleFieldList(FieldList):
    def __init__(self, fields: Iterable[Any] | None = None):
        self._fields = list(fields) if fields is not None else []

    def append(self, field: Any) -> None:
        self._fields.append(field)

    def __len__(self) -> int:
        return len(self._fields)

    def __iter__(self) -> Iterator[Any]:
        return iter(self._fields)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return SimpleFieldList(self._fields[i])
        return self._fields[i]

    def metadata(self, *keys: str, **kwargs: Any) -> list[Any]:
        return [f.metadata(*keys, **kwargs) for f in self._fields]

    def sel(self, **kwargs: Any) -> "SimpleFieldList":
        def ok(f):
            for k, want in kwargs.items():
                try:
                    v = f.metadata(k)
                except KeyError:
                    return False
                if isinstance(want, (list, tuple, set)):
                    if v not in want:
                        return False
                elif v != want:
                    return False
            return True

        return SimpleFieldList([f for f in self._fields if ok(f)])

    def to_numpy(self, **kwargs: Any) -> np.ndarray:
        return np.stack([f.to_numpy(**kwargs) for f in self._fields])

    def __repr__(self) -> str:
        return f"SimpleFieldList({len(self._fields)} fields)"


def _field_from_dict(d: dict[str, Any], mars: bool = False) -> ArrayField:
    values = np.asarray(d["values"])
    lat = d.get("latitudes")
    lon = d.get("longitudes")
    lats = lons = None
    if lat is not None and lon is not None:
        lat, lon = np.asarray(lat, dtype=np.float64), np.asarray(lon, dtype=np.float64)
        if lat.size * lon.size == values.size and not (lat.size == values.size and lon.size == values.size):
            # 1-D axes of a regular grid -> full coordinates, C order
            lats, lons = (a.reshape(-1) for a in np.meshgrid(lat, lon, indexing="ij"))
        else:
            lats, lons = lat.reshape(-1), lon.reshape(-1)
    md = {k: v for k, v in d.items() if k != "values"}
    return ArrayField(values, md, latitudes=lats, longitudes=lons, mars=mars)


def from_source(kind: str, *args: Any, **kwargs: Any) -> SimpleFieldList:
    """`earthkit.data.from_source` for the one in-memory source the path needs."""
    if kind == "list-of-dicts":
        (items,) = args
        return SimpleFieldList([_field_from_dict(d, mars=kwargs.get("mars", False)) for d in items])
    raise NotImplementedError(f"from_source({kind!r}) needs earthkit-data; only 'list-of-dicts' is built in")
