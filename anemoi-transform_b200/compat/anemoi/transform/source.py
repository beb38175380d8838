from anemoi_transform_b200.source import *  # noqa: F401,F403
from anemoi_transform_b200.source import Source  # noqa: F401
