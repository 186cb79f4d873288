#!/usr/bin/env python
"""Golden fixtures of the REAL reference for merges of more than 16 task vectors (the wide path of this build: mask pack
+ blocked / tensor-core Gram + runtime-N pass 2).  Same procedure and same container-only rule as make_golden.py, whose
case runner it reuses; written to a file of its own so that the existing fixtures stay byte-for-byte what they were.

    mkdir -p /tmp/stubs && : > /tmp/stubs/open_clip.py
    PYTHONPATH=/root/reference:/tmp/stubs PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_wide.py
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from make_golden import pipeline_case  # noqa: E402  (imports the reference from PYTHONPATH and asserts its origin)

WIDE_CASES = {
    "wide20_union_uniform": dict(n=20, mask_p=0.3, cfg=dict(svd_mask_strategy="union", svd_energy_threshold=0.9,
                                                            svd_rtvq_stages=2, svd_low_bits=4)),
    "wide24_majority_cluster": dict(n=24, mask_p=0.5, cfg=dict(svd_mask_strategy="majority", svd_weighting="cluster",
                                                               svd_cluster_k=2, svd_energy_threshold=0.95,
                                                               svd_rtvq_stages=2, svd_low_bits=4)),
}


def main():
    torch.manual_seed(0)
    pipe = {name: pipeline_case(name, spec) for name, spec in WIDE_CASES.items()}
    torch.save(pipe, os.path.join(HERE, "pipeline_golden_wide.pt"))
    meta_path = os.path.join(HERE, "golden_meta.json")
    meta = json.load(open(meta_path))
    meta.setdefault("generators", {})["pipeline_golden_wide.pt"] = "tests/golden/make_golden_wide.py"
    meta["cases"].update({k: v["files"] for k, v in pipe.items()})
    with open(meta_path, "w") as f:
        json.dump(meta, f, indent=1)
    print("pipeline_golden_wide.pt", os.path.getsize(os.path.join(HERE, "pipeline_golden_wide.pt")),
          {k: {p: b["masked"]["k"] for p, b in v["bases"].items()} for k, v in pipe.items()})


if __name__ == "__main__":
    main()
