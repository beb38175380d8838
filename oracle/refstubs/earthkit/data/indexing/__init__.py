from . import fieldlist  # noqa: F401
