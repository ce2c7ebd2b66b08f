"""B200-native DRCT / DRN super-resolution + anomaly scoring hot path.

The directory name mirrors the reference repository and is not a valid Python identifier; import it
with `importlib.import_module("anomaly-detection-super-resolution_b200")` (tests/conftest.py and the
root-level `src/` drop-in package do exactly that).  Sub-modules:

  _abi     ctypes binding of include/adsr_b200.h (libadsr_b200.so; no fallback)
  ops      one torch-tensor wrapper per C entry point
  pack     load-time packing of weights into tcgen05 shared-memory images
  drct     DRCT(opt)   -- drop-in for the reference's src.drct.DRCT
  metrics  psnr/ssim   -- drop-in for src.metrics, plus the batched GPU scorer
  evaluate batched evaluator -- drop-in for src.evaluate.evaluate_on_test / CLI
"""
from . import _abi  # noqa: F401

__all__ = ["_abi", "ops", "pack", "drct"]
