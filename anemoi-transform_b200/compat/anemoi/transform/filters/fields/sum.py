from anemoi_transform_b200.filters.fields.sum import *  # noqa: F401,F403
