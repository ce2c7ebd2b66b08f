"""Alias of anomaly-detection-super-resolution_b200/evaluate.py (see src/__init__.py)."""
from . import _alias

_mod = _alias("evaluate")
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})

if __name__ == "__main__" and hasattr(_mod, "main"):
    _mod.main()
