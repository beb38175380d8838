from anemoi_transform_b200.filters.fields.cos_sin_mean_wave_direction import *  # noqa: F401,F403
