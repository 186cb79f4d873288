"""Task-vector containers of the reference's root ``task_vectors.py:61-1012``.

Same class names, constructor arguments, attributes and methods.  Whole-tensor quantisation goes through
``quantization_utils`` (K4 kernels: one global min/max per tensor, bit-exact codes); the element-wise
state-dict arithmetic (ft - base, +, -, * scalar, base + delta) stays plain tensor arithmetic on the
tensors' own device, as in the reference."""
from typing import Dict, Optional, Union

import torch

from . import quantization_utils

Checkpoint = Union[str, Dict[str, torch.Tensor], torch.nn.Module]


def _extract_state_dict(checkpoint) -> Dict[str, torch.Tensor]:
    """task_vectors.py:61-95: nn.Module / {"state_dict"|"model"|"model_state_dict"} wrappers."""
    if isinstance(checkpoint, torch.nn.Module):
        return checkpoint.state_dict()
    if isinstance(checkpoint, dict):
        for key in ("state_dict", "model", "model_state_dict"):
            if key in checkpoint:
                return checkpoint[key]
    return checkpoint


def _load(checkpoint: Checkpoint) -> Dict[str, torch.Tensor]:
    if isinstance(checkpoint, str):
        checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=False)
    return _extract_state_dict(checkpoint)


def _skip(param: torch.Tensor, skip_int64: bool, skip_uint8: bool) -> bool:
    return (skip_int64 and param.dtype == torch.int64) or (skip_uint8 and param.dtype == torch.uint8)


class TaskVector:
    """task_vector = fine-tuned - pretrained, key by key (task_vectors.py:98-635)."""

    def __init__(self, pretrained_checkpoint: Checkpoint, finetuned_checkpoint: Checkpoint,
                 task_name: Optional[str] = None, skip_int64: bool = True, skip_uint8: bool = True,
                 verbose: bool = True):
        self.task_name = task_name
        self.verbose = verbose
        pre, fin = _load(pretrained_checkpoint), _load(finetuned_checkpoint)
        self.vector: Dict[str, torch.Tensor] = {}
        for key, p in pre.items():
            f = fin.get(key)
            if f is None or _skip(p, skip_int64, skip_uint8) or p.shape != f.shape:
                continue
            self.vector[key] = f - p
        if verbose:
            total = sum(d.float().norm().item() ** 2 for d in self.vector.values()) ** 0.5
            print(f"   task vector{' ' + task_name if task_name else ''}: {len(self.vector)} tensors, L2 norm {total:.6f}")

    @classmethod
    def _from_vector(cls, vector, task_name=None):
        out = cls.__new__(cls)
        out.task_name, out.vector, out.verbose = task_name, vector, False
        return out

    def _name(self, other, op):
        return f"{self.task_name}{op}{other.task_name}" if self.task_name and other.task_name else None

    def __add__(self, other: "TaskVector") -> "TaskVector":
        return self._from_vector({k: v + other.vector[k] for k, v in self.vector.items() if k in other.vector},
                                 self._name(other, "+"))

    def __sub__(self, other: "TaskVector") -> "TaskVector":
        return self._from_vector({k: v - other.vector[k] for k, v in self.vector.items() if k in other.vector},
                                 self._name(other, "-"))

    def __mul__(self, scalar: float) -> "TaskVector":
        return self._from_vector({k: v * scalar for k, v in self.vector.items()}, self.task_name)

    def __rmul__(self, scalar: float) -> "TaskVector":
        return self.__mul__(scalar)

    def apply_to(self, pretrained_checkpoint: Checkpoint, verbose: bool = None) -> Dict[str, torch.Tensor]:
        pre = _load(pretrained_checkpoint)
        return {k: (v + self.vector[k] if k in self.vector else (v.clone() if isinstance(v, torch.Tensor) else v))
                for k, v in pre.items()}


def _quantize(param: torch.Tensor, qbit: int, method: str) -> Dict:
    if method == "asymmetric":
        q, scale, zp = quantization_utils.asymmetric_quantization(param, qbit)
        return {"quantized": q, "scale": scale, "zero_point": zp, "shape": param.shape}
    q, scale = quantization_utils.absmax_quantization(param, qbit)
    return {"quantized": q, "scale": scale, "shape": param.shape}


def _dequantize(payload: Dict, method: str) -> torch.Tensor:
    if method == "asymmetric":
        return quantization_utils.dequantize_asymmetric(payload["quantized"], payload["scale"],
                                                        payload.get("zero_point", torch.tensor(0.0)))
    return quantization_utils.dequantize_absmax(payload["quantized"], payload["scale"])


class QuantizedTaskVector:
    """task_vectors.py:638-761: {key: {"quantized", "scale"[, "zero_point"]}} -> deltas."""

    def __init__(self, quantized_deltas: Dict[str, Dict], method: str = "asymmetric"):
        self.quantized_deltas = quantized_deltas
        self.method = method

    def dequantize(self) -> Dict[str, torch.Tensor]:
        return {k: _dequantize(p, self.method) for k, p in self.quantized_deltas.items()}

    def apply_to(self, pretrained_checkpoint: Checkpoint) -> Dict[str, torch.Tensor]:
        pre, vec = _load(pretrained_checkpoint), self.dequantize()
        return {k: (v + vec[k] if k in vec else v) for k, v in pre.items()}


class QuantizedFinetunedModel:
    """task_vectors.py:764-874: every tensor of a fine-tuned checkpoint quantised whole."""

    def __init__(self, finetuned_checkpoint: Checkpoint, qbit: int = 8, method: str = "asymmetric",
                 skip_int64: bool = True, skip_uint8: bool = True):
        self.qbit, self.method = qbit, method
        self.quantized_weights = {k: _quantize(p, qbit, method) for k, p in _load(finetuned_checkpoint).items()
                                  if not _skip(p, skip_int64, skip_uint8)}

    def dequantize(self) -> Dict[str, torch.Tensor]:
        return {k: _dequantize(p, self.method).reshape(p["shape"]) for k, p in self.quantized_weights.items()}

    def get_task_vector(self, pretrained_checkpoint: Checkpoint) -> Dict[str, torch.Tensor]:
        pre = _load(pretrained_checkpoint)
        return {k: v - pre[k] for k, v in self.dequantize().items() if k in pre}


class QuantizedBaseAndTaskVector:
    """task_vectors.py:877-1012: base and task vector quantised separately (RTVQ-style storage)."""

    def __init__(self, pretrained_checkpoint: Checkpoint, task_vector: Union[TaskVector, Dict[str, torch.Tensor]],
                 base_qbit: int = 8, task_qbit: int = 8, method: str = "asymmetric", skip_int64: bool = True,
                 skip_uint8: bool = True):
        self.method, self.base_qbit, self.task_qbit = method, base_qbit, task_qbit
        deltas = task_vector.vector if isinstance(task_vector, TaskVector) else task_vector
        self.quantized_base = {k: _quantize(p, base_qbit, method) for k, p in _load(pretrained_checkpoint).items()
                               if not _skip(p, skip_int64, skip_uint8)}
        self.quantized_task = {k: _quantize(d, task_qbit, method) for k, d in deltas.items()}

    def dequantize(self) -> Dict[str, torch.Tensor]:
        out = {k: _dequantize(p, self.method).reshape(p["shape"]) for k, p in self.quantized_base.items()}
        for k, p in self.quantized_task.items():
            d = _dequantize(p, self.method).reshape(p["shape"])
            out[k] = out[k] + d if k in out else d
        return out
