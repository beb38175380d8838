"""anemoi_transform_b200 — the regrid / spatial / pointwise field-transform hot path of
ecmwf/anemoi-transform, rebuilt for NVIDIA B200 (sm_100a).

Host code is Python (this package) calling libat_b200.so through ctypes; there is no CPU
fallback.  Public surface mirrors the reference:

    from anemoi_transform_b200.filters import filter_registry, create_filter, create_filter_by_name
    from anemoi_transform_b200 import spatial
"""

__version__ = "0.1.0"
