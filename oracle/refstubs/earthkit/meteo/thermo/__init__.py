from oracle.pointwise import dewpoint_from_relative_humidity, relative_humidity_from_dewpoint  # noqa: F401

from . import array  # noqa: F401
