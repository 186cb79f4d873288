"""``python src/main.py --method svd_hybrid ...`` -- same dispatcher contract as the reference's
src/main.py:41-89 (only svd_hybrid is implemented there as well)."""
import argparse
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


def main():
    p = argparse.ArgumentParser(description="Task merging methods for multi-task models")
    p.add_argument("--method", type=str, default="svd_hybrid", choices=["svd_hybrid", "task_arithmetic", "ties", "dare"],
                   help="Merging method to use (default: svd_hybrid)")
    args, rest = p.parse_known_args()
    if args.method != "svd_hybrid":
        print(f"Method {args.method} is not implemented (the reference only implements svd_hybrid)")
        return None
    from svd_quantization_task_merging_b200.svd_hybrid.cli import main as svd_main
    sys.argv = [sys.argv[0]] + rest
    return svd_main()


if __name__ == "__main__":
    main()
