from anemoi_transform_b200.transform import *  # noqa: F401,F403
from anemoi_transform_b200.transform import ReversedTransform, Transform  # noqa: F401
