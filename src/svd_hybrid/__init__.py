"""Import-path alias: ``from src.svd_hybrid.<module> import ...`` (the path the reference's tests,
scripts and load_and_merge.py use) resolves to svd_quantization_task_merging_b200.svd_hybrid."""
import importlib
import sys

_PKG = "svd_quantization_task_merging_b200.svd_hybrid"
_MODULES = ("config", "task_vector_loader", "mask_loader", "basis", "compress", "rtvq", "weighting", "clustering",
            "merge", "diagnostics", "storage", "cli", "run", "reload")
for _m in _MODULES:
    sys.modules[f"{__name__}.{_m}"] = importlib.import_module(f"{_PKG}.{_m}")
    globals()[_m] = sys.modules[f"{__name__}.{_m}"]

from svd_quantization_task_merging_b200.svd_hybrid import SVDHybridConfig, run_svd_hybrid  # noqa: E402,F401
