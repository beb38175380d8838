from . import fieldlist, geography  # noqa: F401
from . import metadata  # noqa: F401,E402
