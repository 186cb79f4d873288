"""Root-level alias of the reference's ``task_vectors`` module (same class names)."""
from svd_quantization_task_merging_b200.task_vectors import (  # noqa: F401
    QuantizedBaseAndTaskVector, QuantizedFinetunedModel, QuantizedTaskVector, TaskVector, _extract_state_dict)
