from anemoi_transform_b200.filters.fields.uv_to_ddff import *  # noqa: F401,F403
