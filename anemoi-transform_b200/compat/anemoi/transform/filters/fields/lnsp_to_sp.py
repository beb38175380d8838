from anemoi_transform_b200.filters.fields.lnsp_to_sp import *  # noqa: F401,F403
