"""Constants of the spatial path (reference `constants.py:11-25`; values are
earthkit.meteo.constants.constants, which is not vendored in the reference)."""

import math

R_earth_meter = 6371229.0
R_earth_km = R_earth_meter / 1000
radian = math.pi / 180.0
L_1_degree_earth_arc_length_km = R_earth_km * radian
g_gravitational_acceleration = 9.80665  # earthkit.meteo.constants.constants.g (reference constants.py)
