"""Host-side mirror of the reference package ``src/svd_hybrid`` (same module names, function names,
argument order, defaults and return structures) over the sm_100a kernels in libsvdq.so.
The evaluation side system of the reference (eval.py, heads.py, dataset/*, src/modeling.py) is out
of scope and is not imported, so no ``open_clip`` is needed to merge."""
from .config import SVDHybridConfig
from .run import run_svd_hybrid

__all__ = ["SVDHybridConfig", "run_svd_hybrid"]
