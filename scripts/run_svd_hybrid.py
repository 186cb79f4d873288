#!/usr/bin/env python3
"""Entry script of the SVD-Hybrid merge (same role and flags as the reference's scripts/run_svd_hybrid.py:1-14):
forwards to the pipeline CLI, which runs the fused B200 path."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from src.svd_hybrid.cli import main  # noqa: E402

if __name__ == "__main__":
    main()
