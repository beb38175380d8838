from oracle.pointwise import (  # noqa: F401
    relative_humidity_from_specific_humidity,
    saturation_vapour_pressure,
    specific_humidity_from_relative_humidity,
)
