from anemoi_transform_b200.filters.fields.rescale import *  # noqa: F401,F403
