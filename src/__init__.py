"""Drop-in `src` package: the reference's import paths (`src.drct`, `src.drn`, `src.model`, `src.metrics`,
`src.evaluate`, `src.main`) resolve to the B200-native implementation in
`anomaly-detection-super-resolution_b200/`, so `python -m src.evaluate ...` keeps working unchanged."""
import importlib as _importlib
import sys as _sys

_IMPL = "anomaly-detection-super-resolution_b200"


def _alias(name: str):
    mod = _importlib.import_module(f"{_IMPL}.{name}")
    _sys.modules[f"{__name__}.{name}"] = mod
    return mod
