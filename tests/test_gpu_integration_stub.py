"""The ctypes stub printed in INTEGRATION.md is executable documentation: run it verbatim against
libsvdq.so and require bit-identical output to the packaged engine."""
import os
import re

import pytest
import torch

from svd_quantization_task_merging_b200 import _native, synth
from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_stub_runs_and_matches_engine(cuda_device):
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(.*?)```", text, re.S).group(1).replace('"libsvdq.so"', repr(_native.LIB_PATH))
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    tasks = synth.task_names(8)
    shapes = {"a.weight": (300, 70), "b.bias": (4099,), "c": (17, 33, 5)}
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=2, device="cuda")
    masks = synth.make_masks(shapes, tasks, 0.5, seed=3, device="cuda")
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy="majority",
                          svd_store_artifacts=False, svd_eval_reconstruction=False)
    w = {t: 1.0 / len(tasks) for t in tasks}
    out, tables = ns["merge_on_gpu"](base, [fts[t] for t in tasks], [masks[t] for t in tasks], cfg, w)
    torch.cuda.synchronize()
    # the stub is the minimal K1 -> K2 -> K3 chain (closed-form coefficients everywhere); the engine's default
    # additionally re-projects small parameters (K7), so compare against projection="closed"
    res = merge_state_dicts(base, fts, masks, cfg, "cuda", projection="closed")
    for n in shapes:
        assert torch.equal(out[n], res["merged_state_dict"][n]), n
    assert (tables["info"][:, 0] == 0).all()
