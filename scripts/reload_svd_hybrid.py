#!/usr/bin/env python3
"""Rebuild a merged model from stored SVD-Hybrid artifacts, optionally checking it against the merged model that was
saved by the original run.  Same command line, helper names and exit codes as the reference's
scripts/reload_svd_hybrid.py:1-256; the reconstruction itself is the batched reload merge on the GPU
(src.svd_hybrid.reload -> K11, one launch over all stored bases)."""
import argparse
import hashlib
import json
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from src.svd_hybrid.reload import reload_merged_model_from_artifacts  # noqa: E402
from src.svd_hybrid.storage import load_merged_model  # noqa: E402

_RULE = "=" * 60


def _banner(title: str):
    print(f"\n{_RULE}\n{title}\n{_RULE}")


def compute_state_dict_checksum(state_dict: dict) -> str:
    """MD5 over the raw bytes of every tensor, keys in sorted order (reload_svd_hybrid.py:22-43)."""
    md5 = hashlib.md5()
    for key in sorted(state_dict):
        value = state_dict[key]
        if isinstance(value, torch.Tensor):
            md5.update(value.cpu().numpy().tobytes())
    return md5.hexdigest()


def verify_reconstruction(artifact_dir: str, merged_model_path: str, verbose: bool = True) -> bool:
    """True when the model rebuilt from ``artifact_dir`` is byte-identical to the saved merged model
    (reload_svd_hybrid.py:46-115).  With ``verbose`` a mismatch is broken down per parameter."""
    _banner("Verifying Reconstruction")
    print(f"\nLoading saved merged model from {merged_model_path}")
    saved = load_merged_model(merged_model_path)
    saved_sum = compute_state_dict_checksum(saved)
    print(f"Reloading model from artifacts in {artifact_dir}")
    reloaded = reload_merged_model_from_artifacts(artifact_dir)
    reloaded_sum = compute_state_dict_checksum(reloaded)
    same = saved_sum == reloaded_sum
    if verbose:
        print(f"\nSaved model checksum:    {saved_sum}")
        print(f"Reloaded model checksum: {reloaded_sum}")
        print(f"Checksums match: {same}")
        if not same:
            print("\nParameter-wise comparison:")
            for key in sorted(set(saved) | set(reloaded)):
                if key not in saved:
                    print(f"  {key}: MISSING in saved model")
                elif key not in reloaded:
                    print(f"  {key}: MISSING in reloaded model")
                elif saved[key].shape != reloaded[key].shape:
                    print(f"  {key}: SHAPE MISMATCH ({saved[key].shape} vs {reloaded[key].shape})")
                else:
                    gap = (saved[key].float() - reloaded[key].float()).abs()
                    if gap.max().item() > 1e-5:
                        print(f"  {key}: DIFF max={gap.max().item():.6e}, mean={gap.mean().item():.6e}")
    return same


def run_evaluation(merged_model_path: str, tasks: list, eval_script: str = None) -> dict:
    """Placeholder hook, as in the reference (reload_svd_hybrid.py:118-146): evaluation lives outside this script."""
    _banner("Evaluation")
    if eval_script is None:
        print("No evaluation script specified, skipping evaluation")
        return {}
    print(f"Would evaluate {merged_model_path} on tasks: {tasks}")
    print("Evaluation not implemented in this script")
    return {}


def main() -> int:
    ap = argparse.ArgumentParser(description="Reload SVD-Hybrid merged model from artifacts")
    ap.add_argument("--artifact-dir", type=str, required=True, help="Directory containing saved artifacts")
    ap.add_argument("--output-path", type=str, default=None, help="Path to save reloaded merged model (optional)")
    ap.add_argument("--verify", action="store_true", help="Verify against original merged model")
    ap.add_argument("--merged-model-path", type=str, default=None, help="Path to original merged model for verification")
    ap.add_argument("--eval", action="store_true", help="Run evaluation after reloading")
    ap.add_argument("--eval-script", type=str, default=None, help="Path to evaluation script")
    ap.add_argument("--verbose", action="store_true", help="Print detailed information")
    args = ap.parse_args()

    if not os.path.exists(args.artifact_dir):
        print(f"Error: Artifact directory not found: {args.artifact_dir}")
        return 1
    config = {}
    config_path = os.path.join(args.artifact_dir, "config.json")
    if os.path.exists(config_path):
        with open(config_path) as f:
            config = json.load(f)
        if args.verbose:
            print("\nLoaded configuration:")
            print(json.dumps(config, indent=2))
    else:
        print(f"Warning: No config.json found in {args.artifact_dir}")

    _banner("Reloading Model from Artifacts")
    print(f"Artifact directory: {args.artifact_dir}")
    try:
        merged = reload_merged_model_from_artifacts(args.artifact_dir)
    except Exception as exc:  # noqa: BLE001 -- the reference reports the error and exits 1
        print(f"\nError reloading model: {exc}")
        traceback.print_exc()
        return 1
    print(f"\nSuccessfully reloaded model with {len(merged)} parameters")
    if args.verbose:
        tensors = {k: v for k, v in merged.items() if isinstance(v, torch.Tensor)}
        print(f"Total parameters: {sum(v.numel() for v in tensors.values()):,}")
        print("\nParameter shapes:")
        for key in sorted(tensors)[:10]:
            print(f"  {key}: {tensors[key].shape}")
        if len(merged) > 10:
            print(f"  ... and {len(merged) - 10} more")

    if args.output_path:
        print(f"\nSaving reloaded model to {args.output_path}")
        os.makedirs(os.path.dirname(args.output_path) or ".", exist_ok=True)
        torch.save(merged, args.output_path)
        print("Model saved successfully")

    if args.verify:
        if not args.merged_model_path:
            print("\nError: --merged-model-path required for verification")
            return 1
        if not os.path.exists(args.merged_model_path):
            print(f"\nError: Merged model not found: {args.merged_model_path}")
            return 1
        if verify_reconstruction(args.artifact_dir, args.merged_model_path, verbose=args.verbose):
            print("\n✓ Verification PASSED: Reloaded model matches original")
        else:
            print("\n✗ Verification FAILED: Reloaded model differs from original")
            return 1

    if args.eval:
        results = run_evaluation(args.output_path or "reloaded_model.pt", config.get("tasks", []), args.eval_script)
        if results:
            print("\nEvaluation Results:")
            print(json.dumps(results, indent=2))

    _banner("Reload Complete!")
    return 0


if __name__ == "__main__":
    sys.exit(main())
