from anemoi_transform_b200.filters.fields.apply_mask import *  # noqa: F401,F403
