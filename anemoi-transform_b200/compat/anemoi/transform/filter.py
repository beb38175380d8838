from anemoi_transform_b200.filter import *  # noqa: F401,F403
from anemoi_transform_b200.filter import DispatchingFilter, Filter, SingleFieldFilter  # noqa: F401
