"""Alias of anomaly-detection-super-resolution_b200/training.py (see src/__init__.py)."""
from . import _alias

_mod = _alias("training")
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
