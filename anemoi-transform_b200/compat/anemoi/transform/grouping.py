from anemoi_transform_b200.grouping import *  # noqa: F401,F403
from anemoi_transform_b200.grouping import GroupByParam  # noqa: F401
