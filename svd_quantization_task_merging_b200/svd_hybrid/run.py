"""Programmatic entry.  Mirror of src/svd_hybrid/run.py:40-91."""
from typing import Dict

from .cli import run_svd_hybrid_pipeline
from .config import SVDHybridConfig


def run_svd_hybrid(config: SVDHybridConfig) -> Dict:
    return run_svd_hybrid_pipeline(config)


def main(args=None):
    from .cli import main as cli_main
    return cli_main()


if __name__ == "__main__":
    main()
