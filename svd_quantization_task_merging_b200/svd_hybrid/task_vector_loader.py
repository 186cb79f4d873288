"""Checkpoint ingest and task-vector construction.  Mirror of src/svd_hybrid/task_vector_loader.py:56-291.

The fused path (engine.MergeJob) never materialises task vectors: K1/K3 form ``ft - base`` in
registers.  These functions keep the reference's signatures for callers that want the vectors."""
import os
from typing import Dict, List, Optional, Tuple

import torch

_CKPT_KEYS = ("state_dict", "model", "model_state_dict")


def load_checkpoint(checkpoint_path: str, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """task_vector_loader.py:56-100: torch.load, unwrap nn.Module / {"state_dict"|"model"|"model_state_dict"}."""
    if not os.path.exists(checkpoint_path):
        raise FileNotFoundError(f"Checkpoint not found: {checkpoint_path}")
    if checkpoint_path.endswith(".safetensors"):
        # additive input format (the reference reads pickles only): memory-mapped tensors, no unpickling pass --
        # engine.upload_state_dicts then packs them straight from the page cache into its pinned staging buffers
        from safetensors.torch import load_file
        return load_file(checkpoint_path, device=device)
    ckpt = torch.load(checkpoint_path, map_location=device, weights_only=False)
    if isinstance(ckpt, torch.nn.Module):
        return ckpt.state_dict()
    if isinstance(ckpt, dict):
        for key in _CKPT_KEYS:
            if key in ckpt:
                return ckpt[key]
    return ckpt


def compute_task_vector(base_state: Dict[str, torch.Tensor], finetuned_state: Dict[str, torch.Tensor],
                        device: str = "cpu") -> Dict[str, torch.Tensor]:
    """task_vector_loader.py:103-145: delta[k] = ft[k] - base[k] for keys in both with equal shape."""
    out = {}
    for key, b in base_state.items():
        f = finetuned_state.get(key)
        if f is None:
            continue
        b, f = b.to(device), f.to(device)
        if b.shape != f.shape:
            print(f"Warning: Shape mismatch for {key}, skipping")
            continue
        out[key] = (f - b).detach()
    return out


def load_task_vectors(base_model_path: str, task_checkpoint_paths: Dict[str, str], device: str = "cpu",
                      filter_keys: Optional[List[str]] = None) -> Dict[str, Dict[str, torch.Tensor]]:
    def keep(sd):
        return sd if filter_keys is None else {k: v for k, v in sd.items() if any(pat in k for pat in filter_keys)}
    base = keep(load_checkpoint(base_model_path, device))
    return {task: compute_task_vector(base, keep(load_checkpoint(path, device)), device)
            for task, path in task_checkpoint_paths.items()}


def get_parameter_names(task_vectors: Dict[str, Dict[str, torch.Tensor]]) -> List[str]:
    return sorted({p for tv in task_vectors.values() for p in tv})


def organize_by_parameter(task_vectors: Dict[str, Dict[str, torch.Tensor]]) -> Dict[str, Dict[str, torch.Tensor]]:
    out: Dict[str, Dict[str, torch.Tensor]] = {}
    for task, tv in task_vectors.items():
        for p, d in tv.items():
            out.setdefault(p, {})[task] = d
    return out


def flatten_task_deltas(task_vectors: Dict[str, Dict[str, torch.Tensor]], param_name: str
                        ) -> Tuple[List[torch.Tensor], List[str]]:
    deltas, names = [], []
    for task, tv in task_vectors.items():
        if param_name in tv:
            deltas.append(tv[param_name].flatten())
            names.append(task)
    return deltas, names


def get_task_checkpoint_paths(checkpoint_dir: str, task_names: List[str]) -> Dict[str, str]:
    """task_vector_loader.py:258-291: {t}.pt, {t}.pth, {t}/checkpoint.pt, {t}/model.pt, {t}/finetuned.pt; then, as an
    additive format, {t}.safetensors and {t}/model.safetensors."""
    out = {}
    for t in task_names:
        cands = [os.path.join(checkpoint_dir, f"{t}.pt"), os.path.join(checkpoint_dir, f"{t}.pth"),
                 os.path.join(checkpoint_dir, t, "checkpoint.pt"), os.path.join(checkpoint_dir, t, "model.pt"),
                 os.path.join(checkpoint_dir, t, "finetuned.pt"),
                 os.path.join(checkpoint_dir, f"{t}.safetensors"),                 # additive, tried last
                 os.path.join(checkpoint_dir, t, "model.safetensors")]
        hit = next((c for c in cands if os.path.exists(c)), None)
        if hit is None:
            raise FileNotFoundError(f"No checkpoint found for task {t} in {checkpoint_dir}")
        out[t] = hit
    return out
