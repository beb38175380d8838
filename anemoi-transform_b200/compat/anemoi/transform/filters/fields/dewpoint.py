from anemoi_transform_b200.filters.fields.dewpoint import *  # noqa: F401,F403
