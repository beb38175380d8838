"""earthkit.meteo.constants.constants — the four names reference constants.py:11-14 imports."""
import numpy as np

R = 8.31446261815324
R_earth = 6371229.0
g = 9.80665
radian = np.pi / 180.0
