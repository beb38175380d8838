from anemoi_transform_b200.filters.fields.regrid import *  # noqa: F401,F403
