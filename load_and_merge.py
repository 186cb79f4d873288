"""``python load_and_merge.py --artifact-dir A --base-model-path B --output-path O`` -- rebuild a merged
model from stored SVD-Hybrid artifacts (the reference's root ``load_and_merge.py:20-127``).  Thin wrapper
over ``svd_hybrid.reload``; the re-merge runs on the GPU, results are returned on ``--device``."""
import argparse

from svd_quantization_task_merging_b200.svd_hybrid.reload import reconstruct_from_artifacts as _reconstruct


def reconstruct_from_artifacts(artifact_dir: str, base_model_path: str, output_path: str, device=None):
    return _reconstruct(artifact_dir, base_model_path, output_path, "cpu" if device in (None, "auto") else device)


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
