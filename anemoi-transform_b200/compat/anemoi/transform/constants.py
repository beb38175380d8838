from anemoi_transform_b200.constants import *  # noqa: F401,F403
