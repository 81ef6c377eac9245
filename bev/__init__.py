"""Drop-in import name of the reference package: ``from bev import BEV, bev_settings`` (simple_romp/bev/__init__.py)
resolves to the B200-native implementation in ``romp_b200.bev``; ``bev.main.default_settings`` is lazy."""
from romp_b200.bev import BEV, bev_settings  # noqa: F401
from . import main  # noqa: F401

__all__ = ["BEV", "bev_settings", "main"]
