from anemoi_transform_b200.filters.fields.clipper import *  # noqa: F401,F403
