"""Stand-in for libpysal==4.4.0 (absent) -- TEST INFRASTRUCTURE.  Only ``weights.lat2SW`` (utils.py:98)."""
from . import weights  # noqa: F401
