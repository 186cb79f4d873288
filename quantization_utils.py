"""Root-level alias of the reference's ``quantization_utils`` module (same function names)."""
from svd_quantization_task_merging_b200.quantization_utils import *  # noqa: F401,F403
from svd_quantization_task_merging_b200.quantization_utils import (  # noqa: F401
    absmax_quantization, asymmetric_quantization, dequantize_absmax, dequantize_asymmetric,
    quantization_error_check_asymmetric, qunatization_error_check)
