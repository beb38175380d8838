from . import fieldlist, geography  # noqa: F401
