"""A name → factory registry with the API the reference uses from `anemoi.utils.registry`.

Call sites mirrored (reference src/anemoi/transform/): `filters/__init__.py:25-31,38,58`,
`transform.py:131`, `commands/filters.py:45`, `tests/test_create.py:19-21`:
`Registry(package)`, `.register(name, factory=None, aliases=None)` (decorator when the
factory is omitted), `.create(name, *args, **kwargs)`, `.from_config(config)`,
`.registered`, `.lookup(name, return_none=False)`, `.factories`, `.aliases()`, `.package`,
`.is_registered(name)`.

Third-party plugins can add filters through the `anemoi.transform.filters` entry-point group
(loaded lazily the first time a name is not found).
"""

from __future__ import annotations

import logging
from typing import Any, Callable

LOG = logging.getLogger(__name__)


class Registry:
    def __init__(self, package: str, key: str = "_type", entry_point_group: str | None = None):
        self.package = package
        self.key = key
        self.entry_point_group = entry_point_group
        self._factories: dict[str, Callable[..., Any]] = {}
        self._aliases: dict[str, list[str]] = {}
        self._alias_of: dict[str, str] = {}
        self._plugins_loaded = False

    # -- registration ------------------------------------------------------------------
    def register(self, name: str, factory: Callable[..., Any] | None = None, aliases: list[str] | None = None):
        if factory is None:

            def decorator(f: Callable[..., Any]) -> Callable[..., Any]:
                self.register(name, f, aliases=aliases)
                return f

            return decorator

        assert name not in self._factories and name not in self._alias_of, f"Duplicate registration of '{name}' in {self.package}"
        self._factories[name] = factory
        for alias in aliases or []:
            assert alias not in self._factories and alias not in self._alias_of, f"Duplicate alias '{alias}' in {self.package}"
            self._alias_of[alias] = name
            self._aliases.setdefault(name, []).append(alias)
        return None

    # -- introspection -----------------------------------------------------------------
    @property
    def factories(self) -> dict[str, Callable[..., Any]]:
        return dict(self._factories)

    def aliases(self) -> dict[str, list[str]]:
        return {k: list(v) for k, v in self._aliases.items()}

    @property
    def registered(self) -> list[str]:
        return sorted(self._factories)

    def is_registered(self, name: str) -> bool:
        return self.lookup(name, return_none=True) is not None

    def _load_plugins(self) -> None:
        if self._plugins_loaded or not self.entry_point_group:
            return
        self._plugins_loaded = True
        try:
            from importlib.metadata import entry_points

            for ep in entry_points(group=self.entry_point_group):
                if ep.name not in self._factories:
                    try:
                        self._factories[ep.name] = ep.load()
                    except Exception as e:  # a broken plugin must not take the registry down
                        LOG.warning("could not load plugin %s: %s", ep.name, e)
        except Exception as e:
            LOG.debug("entry points unavailable: %s", e)

    def lookup(self, name: str, return_none: bool = False) -> Callable[..., Any] | None:
        name = self._alias_of.get(name, name)
        if name not in self._factories:
            self._load_plugins()
        factory = self._factories.get(name)
        if factory is None and not return_none:
            raise ValueError(f"Cannot load '{name}' from {self.package}. Registered: {self.registered}")
        return factory

    # -- creation ----------------------------------------------------------------------
    def create(self, name: str, *args: Any, **kwargs: Any) -> Any:
        return self.lookup(name)(*args, **kwargs)

    def from_config(self, config: Any, *args: Any, **kwargs: Any) -> Any:
        """`"name"`, `{"name": {kwargs}}` / `{"name": value}` or `{"_type": "name", **kwargs}`."""
        if isinstance(config, str):
            return self.create(config, *args, **kwargs)
        if not isinstance(config, dict):
            raise ValueError(f"Invalid config {config!r}: expected a string or a dict")
        if self.key in config:
            config = dict(config)
            name = config.pop(self.key)
            return self.create(name, *args, **config, **kwargs)
        if len(config) == 1:
            ((name, value),) = config.items()
            if isinstance(value, dict):
                return self.create(name, *args, **value, **kwargs)
            if value is None:
                return self.create(name, *args, **kwargs)
            return self.create(name, value, *args, **kwargs)
        raise ValueError(f"Entry '{config}' must either be a string, a dictionary with a single entry, or have a '{self.key}' key")

    def __call__(self, name: str, *args: Any, **kwargs: Any) -> Any:
        return self.create(name, *args, **kwargs)
