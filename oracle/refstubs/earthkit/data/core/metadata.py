"""earthkit.data.core.metadata.RawMetadata — a dict of metadata (the reference's test helper
`tests/utils/__init__.py:mock_field` subclasses it to add a "mars" namespace)."""


class RawMetadata(dict):
    geography = None

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def as_namespace(self, namespace=None):
        return {}

    def override(self, *args, **kwargs):
        d = type(self)(self)
        for a in args:
            d.update(a)
        d.update(kwargs)
        return d
