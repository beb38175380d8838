"""earthkit.data.utils.metadata.dict.UserMetadata — metadata built from a user dict
(`tests/conftest.py:MarsUserMetadata` subclasses it)."""
from earthkit.data.core.metadata import RawMetadata


class UserMetadata(RawMetadata):
    def __init__(self, d=None, shape=None, **kwargs):
        super().__init__({k: v for k, v in dict(d or {}, **kwargs).items() if k != "values"})
        self._data = dict(self)
        self.shape = shape
