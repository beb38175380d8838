"""Stand-in for anemoi.utils.caching (not installed): `cached` without a cache."""


def cached(*args, **kwargs):
    def decorator(func):
        return func

    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    return decorator
