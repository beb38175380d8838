from anemoi_transform_b200.matching import MatchingFieldsFilter, MatchingSpec  # noqa: F401
