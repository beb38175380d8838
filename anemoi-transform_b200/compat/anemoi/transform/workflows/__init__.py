from anemoi_transform_b200.workflows import *  # noqa: F401,F403
