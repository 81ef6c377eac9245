"""``bev.main`` of the reference (simple_romp/bev/main.py): BEV, bev_settings and the lazily evaluated module-level
``default_settings`` (bev/main.py:61)."""
from romp_b200.bev import BEV, bev_settings  # noqa: F401

_default = None


def __getattr__(name):
    global _default
    if name == "default_settings":
        if _default is None:
            _default = bev_settings([])
        return _default
    raise AttributeError(name)
