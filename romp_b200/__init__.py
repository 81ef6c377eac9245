"""romp_b200 - B200 (sm_100a) implementation of ROMP's per-frame inference hot path behind simple_romp's
``ROMP(settings)(image)`` call surface.  ``from romp_b200 import ROMP, romp_settings`` mirrors
``from romp import ROMP, romp_settings`` (simple_romp/romp/__init__.py:1)."""
from .main import ROMP, romp_settings  # noqa: F401

__all__ = ["ROMP", "romp_settings"]
