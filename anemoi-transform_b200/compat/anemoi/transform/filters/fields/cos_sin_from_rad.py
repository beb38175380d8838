from anemoi_transform_b200.filters.fields.cos_sin_from_rad import *  # noqa: F401,F403
