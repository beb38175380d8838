from anemoi_transform_b200.filters.fields.q_to_r import *  # noqa: F401,F403
