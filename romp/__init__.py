"""Drop-in import name of the reference package: ``from romp import ROMP, romp_settings`` (simple_romp/romp/__init__.py:1)
resolves to the B200-native implementation in ``romp_b200``.  Nothing is downloaded or evaluated at import time; the
reference's module-level ``romp.main.default_settings`` (main.py:62) is provided lazily by ``romp.main``."""
from romp_b200.main import ROMP, romp_settings  # noqa: F401
from . import main  # noqa: F401

__all__ = ["ROMP", "romp_settings", "main"]
