from anemoi_transform_b200.filters.fields.impute_nans import *  # noqa: F401,F403
