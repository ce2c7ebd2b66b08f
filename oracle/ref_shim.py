"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (read-only, /root/reference).

Only used in the build container (the reference tree does not travel to the GPU box) by
`oracle/make_golden.py` and by the `-m "not gpu"` tests that pin the oracle to the reference.
Nothing under the product package may import this file.

The reference imports `skimage`, `imageio` and `matplotlib` at module top (src/drct.py:6,11-13,
src/drn.py:6,11-13,26, src/main.py:2,24) although none is used on the model / scoring path.
They are absent in this image, so empty stub modules are registered before the import
(SURVEY.md section 8c).  No reference file is edited or copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ADSR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "drct.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave like a package
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def install_stubs() -> None:
    for name in ("matplotlib", "skimage", "imageio"):
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        if name == "matplotlib":
            _stub("matplotlib", use=lambda *a, **k: None)
            _stub("matplotlib.pyplot")
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        elif name == "skimage":
            _stub("skimage")
            for sub in ("color", "metrics", "exposure"):
                _stub(f"skimage.{sub}")
                setattr(sys.modules["skimage"], sub, sys.modules[f"skimage.{sub}"])

            def _unavailable(*a, **k):  # names imported by src/helpers.py:17-18, never called on this path
                raise RuntimeError("skimage is not installed (stub)")

            _stub("skimage.metrics", structural_similarity=_unavailable, peak_signal_noise_ratio=_unavailable)
        elif name == "imageio":
            import numpy as np
            from PIL import Image

            _stub("imageio")
            _stub("imageio.v2", imread=lambda p: np.array(Image.open(p)))
            sys.modules["imageio"].v2 = sys.modules["imageio.v2"]


def import_reference():
    """Return the reference's `src` package (modules drct, drn, metrics, main, model importable).

    The reference package is literally called `src`; our own drop-in mirror at the repo root has
    the same name, so this must run in a process where the repo root is NOT ahead of the reference
    on sys.path (make_golden.py and the pinning tests take care of that by using a subprocess or
    by inserting the reference root first and purging `src*` from sys.modules).
    """
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        pkg = importlib.import_module("src")
        for sub in ("metrics", "drct", "drn", "main", "model"):
            importlib.import_module(f"src.{sub}")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return pkg


def purge_reference() -> None:
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
