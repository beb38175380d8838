from anemoi_transform_b200.filters.fields.remove_nans import *  # noqa: F401,F403
