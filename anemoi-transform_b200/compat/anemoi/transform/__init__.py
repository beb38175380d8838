"""`anemoi.transform`, served by anemoi_transform_b200 for the field-transform hot path.

Put `anemoi-transform_b200/compat` (and `anemoi-transform_b200`) on `sys.path` and
`import anemoi.transform.spatial`, `anemoi.transform.filters`, … resolve to the B200
implementation: same module paths, names and signatures as the reference for everything
SURVEY.md §8 lists.  `anemoi` itself is a namespace package (no `__init__`), so other
`anemoi.*` distributions (anemoi-utils, …) keep working next to it.
"""

from anemoi_transform_b200 import __doc__ as _doc  # noqa: F401
