"""Rebuild a merged model from stored artifacts.  Mirror of src/svd_hybrid/reload.py:60-267.

Masks are not part of the reference's artifact layout (reload.py:204 passes none, so the reference can only
re-merge unmasked runs).  This build writes the combined masks as one additive file (storage.save_combined_masks)
and uses it here when present; artifacts written by the reference reload exactly as they do there."""
import argparse
import os
from typing import Dict

import torch

from .merge import apply_merged_deltas, merge_all_parameters
from .storage import load_all_artifacts, load_combined_masks
from .task_vector_loader import load_checkpoint
from .weighting import compute_uniform_weights


def _merge_from_artifacts(artifact_dir: str, device: str):
    art = load_all_artifacts(artifact_dir, device=device)
    diag = art["diagnostics"]
    tasks = list(diag.get("task_weights", {}).keys()) or list(next(iter(art["compressed"].values())).keys())
    weights = diag["task_weights"] if "task_weights" in diag else compute_uniform_weights(tasks)
    shapes = {p: torch.Size(d["original_shape"]) for p, d in diag.get("per_parameter", {}).items()
              if d.get("original_shape") is not None}
    masks = load_combined_masks(artifact_dir, device=device)
    deltas = merge_all_parameters(art["compressed"], art["bases"], masks, weights, shapes, art["config"],
                                  device=device, verbose=False)
    return art, deltas


def reload_merged_model_from_artifacts(artifact_dir: str, device: str = "cpu") -> Dict[str, torch.Tensor]:
    path = os.path.join(artifact_dir, "merged_state_dict.pt")
    if os.path.exists(path):
        return torch.load(path, map_location=device, weights_only=False)
    art, deltas = _merge_from_artifacts(artifact_dir, device)
    base_path = getattr(art["config"], "base_model_path", "")      # the reference calls .get() on the dataclass here
    if not base_path or not os.path.exists(base_path):
        raise FileNotFoundError(f"Base model path not found in config or doesn't exist: {base_path}. "
                                "Please provide base model path in artifacts config or use "
                                "reconstruct_from_artifacts instead.")
    return apply_merged_deltas(load_checkpoint(base_path, device=device), deltas, device=device, verbose=False)


def reconstruct_from_artifacts(artifact_dir: str, base_model_path: str, output_path: str, device: str = "cpu") -> Dict:
    art, deltas = _merge_from_artifacts(artifact_dir, device)
    merged = apply_merged_deltas(load_checkpoint(base_model_path, device=device), deltas, device=device, verbose=False)
    torch.save(merged, output_path)
    return {"merged_state_dict": merged, "diagnostics": art["diagnostics"], "config": art["config"]}


def main():
    p = argparse.ArgumentParser(description="Reconstruct merged model from SVD-Hybrid artifacts")
    p.add_argument("--artifact-dir", type=str, required=True, help="Directory containing artifacts")
    p.add_argument("--base-model-path", type=str, required=True, help="Path to base model checkpoint")
    p.add_argument("--output-path", type=str, required=True, help="Path to save reconstructed merged model")
    p.add_argument("--device", type=str, default="cpu", help="Device for the returned tensors")
    a = p.parse_args()
    reconstruct_from_artifacts(a.artifact_dir, a.base_model_path, a.output_path, a.device)


if __name__ == "__main__":
    main()
