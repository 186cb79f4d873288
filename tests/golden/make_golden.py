#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the REAL reference.

Run in the build container only (the reference does not travel to the GPU box):

    mkdir -p /tmp/stubs && : > /tmp/stubs/open_clip.py
    PYTHONPATH=/root/reference:/tmp/stubs PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference package imports ``open_clip`` at import time (src/svd_hybrid/__init__.py:78 ->
eval.py:11 -> src/modeling.py:3); an empty stub module satisfies that import and touches no
hot-path arithmetic.  Nothing is written into /root/reference.  Inputs are stored next to the
outputs so the fixtures are self-contained.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.append(ROOT)          # after PYTHONPATH, so `src` and `quantization_utils` resolve to the reference

from src.svd_hybrid import basis as ref_basis          # noqa: E402  (the reference, from PYTHONPATH)
from src.svd_hybrid import mask_loader as ref_masks    # noqa: E402
from src.svd_hybrid import rtvq as ref_rtvq            # noqa: E402
from src.svd_hybrid.cli import run_svd_hybrid_pipeline  # noqa: E402
from src.svd_hybrid.config import SVDHybridConfig      # noqa: E402
import quantization_utils as ref_qu                    # noqa: E402  (reference root module)

from svd_quantization_task_merging_b200 import synth   # noqa: E402  (only the synthetic generators)

assert "/root/reference" in os.path.abspath(ref_rtvq.__file__), "the reference must come from /root/reference"


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def rtvq_cases():
    out = []
    g = torch.Generator().manual_seed(20261018)
    specs = [(1, 4, 2), (2, 4, 2), (3, 4, 3), (5, 4, 2), (7, 2, 3), (7, 8, 1), (19, 3, 4), (100, 4, 2), (1000, 2, 4),
             (4097, 4, 3), (4097, 8, 2)]
    for n, bits, stages in specs:
        x = torch.randn(n, generator=g) * 0.01
        pay = ref_rtvq.multistage_residual_quantization(x, bits, stages)
        deq = ref_rtvq.multistage_residual_dequantization(pay)
        out.append({"x": x, "bits": bits, "stages": stages, "payloads": pay, "dequantized": deq})
    for x in ([1.0, 2.0, 3.0, 4.0, 5.0], [0.5], [0.0, 0.0, 0.0], [1.0, 1.0], [0.0, 1e-30], [-3.0, 0.0, 2.0]):
        x = torch.tensor(x)
        pay = ref_rtvq.multistage_residual_quantization(x, 4, 2)
        out.append({"x": x, "bits": 4, "stages": 2, "payloads": pay,
                    "dequantized": ref_rtvq.multistage_residual_dequantization(pay)})
    # single-stage quantiser of the root module, 8 and 16 bit, and absmax
    root = []
    for bits in (4, 8, 16):
        x = torch.randn(3000, generator=g) * 3
        q, s, z = ref_qu.asymmetric_quantization(x, bits)
        qa, sa = ref_qu.absmax_quantization(x, bits)
        root.append({"x": x, "bits": bits, "asym": (q, s, z), "asym_deq": ref_qu.dequantize_asymmetric(q, s, z),
                     "absmax": (qa, sa), "absmax_deq": ref_qu.dequantize_absmax(qa, sa)})
    return {"rtvq": out, "root": root}


def mask_cases():
    g = torch.Generator().manual_seed(7)
    out = []
    for n_tasks in (1, 2, 3, 4, 5, 8):
        masks = {f"t{i}": {"w": torch.rand(6, 11, generator=g) < 0.5, "b": torch.rand(33, generator=g) < 0.3}
                 for i in range(n_tasks)}
        if n_tasks >= 3:
            del masks["t1"]["b"]              # a task without a mask for one parameter
        if n_tasks >= 4:
            masks["t2"] = None                # a task without any mask
        case = {"masks": masks}
        for strat in ("union", "intersection", "majority"):
            case[strat] = quiet(ref_masks.combine_masks, masks, strategy=strat, verbose=False)
        out.append(case)
    return out


def rank_cases():
    spectra = [[10.0, 5.0, 2.0, 1.0, 0.5, 0.2, 0.1, 0.05], [100.0, 0.01, 0.001], [10.0, 1e-10, 1e-12], [4.0, 3.0, 2.0, 1.0],
               [10.0, 5.0, 2.0, 1.0, 0.5], [1.0] * 20, [0.0, 0.0], [3.0]]
    g = torch.Generator().manual_seed(3)
    spectra += [torch.sort(torch.rand(n, generator=g), descending=True).values.tolist() for n in (8, 8, 20, 20, 5)]
    out = []
    for sp in spectra:
        S = torch.tensor(sp)
        for thr in (0.5, 0.8, 0.9, 0.95, 0.99, 0.999):
            for mr in (None, 3, 64):
                out.append({"S": S, "thr": thr, "max_rank": mr, "k": ref_basis.select_rank(S, thr, mr),
                            "cum": ref_basis.compute_energy_spectrum(S)})
    out.append({"S": torch.tensor([100.0, 0.01, 0.001]), "thr": 0.999, "max_rank": None, "min_rank": 2,
                "k": ref_basis.select_rank(torch.tensor([100.0, 0.01, 0.001]), 0.999, None, min_rank=2),
                "cum": ref_basis.compute_energy_spectrum(torch.tensor([100.0, 0.01, 0.001]))})
    return out


PIPELINE_CASES = {
    "union_uniform": dict(n=8, mask_p=0.3, cfg=dict(svd_mask_strategy="union", svd_energy_threshold=0.9,
                                                    svd_rtvq_stages=2, svd_low_bits=4)),
    "majority_performance_3stage": dict(n=8, mask_p=0.5, perf=True,
                                        cfg=dict(svd_mask_strategy="majority", svd_weighting="performance",
                                                 svd_weighting_temperature=5.0, svd_energy_threshold=0.9,
                                                 svd_rtvq_stages=3)),
    "intersection_cluster": dict(n=8, mask_p=0.9, cfg=dict(svd_mask_strategy="intersection", svd_weighting="cluster",
                                                           svd_cluster_k=2, svd_energy_threshold=0.9)),
    "nomask_fp32_nocenter": dict(n=6, mask_p=None, cfg=dict(svd_fp16=False, svd_center=False,
                                                            svd_energy_threshold=0.8)),
    "iid_degenerate_nan": dict(n=8, mask_p=None, family="throughput", cfg=dict(svd_energy_threshold=0.95)),
    # noise region (svd_include_noise): a second basis over the rows outside the combined mask
    "majority_noise_uniform": dict(n=8, mask_p=0.5, cfg=dict(svd_mask_strategy="majority", svd_energy_threshold=0.9,
                                                             svd_include_noise=True, svd_noise_shrink=0.5)),
    "union_noise_cluster_3stage": dict(n=6, mask_p=0.3, cfg=dict(svd_mask_strategy="union", svd_weighting="cluster",
                                                                 svd_cluster_k=2, svd_energy_threshold=0.85,
                                                                 svd_rtvq_stages=3, svd_include_noise=True,
                                                                 svd_noise_shrink=0.25)),
}


def pipeline_case(name, spec):
    shapes = synth.toy_shapes()
    tasks = synth.task_names(spec["n"])
    base, fts = synth.make_checkpoints(shapes, tasks, family=spec.get("family", "parity"), seed=1234)
    masks = synth.make_masks(shapes, tasks, spec["mask_p"], seed=99) if spec["mask_p"] is not None else None
    with tempfile.TemporaryDirectory() as d:
        ckpt, mdir, out, art = (os.path.join(d, x) for x in ("ckpt", "masks", "out", "art"))
        for x in (ckpt, mdir, out, art):
            os.makedirs(x)
        torch.save(dict(base), os.path.join(d, "base.pt"))
        for t in tasks:
            torch.save(dict(fts[t]), os.path.join(ckpt, f"{t}.pt"))
            if masks is not None:
                torch.save(dict(masks[t]), os.path.join(mdir, f"{t}_mask.pt"))
        perf_file = None
        if spec.get("perf"):
            perf_file = os.path.join(d, "perf.json")
            with open(perf_file, "w") as f:
                json.dump(synth.performance_table(tasks), f)
        cfg = SVDHybridConfig(tasks=tasks, checkpoint_dir=ckpt, base_model_path=os.path.join(d, "base.pt"),
                              mask_dir=mdir if masks is not None else "", performance_file=perf_file,
                              svd_store_artifacts=True, svd_eval_reconstruction=True, output_dir=out, artifact_dir=art,
                              device="cpu", svd_max_rank=64, **spec["cfg"])
        res = quiet(run_svd_hybrid_pipeline, cfg)
        files = sorted(os.path.relpath(os.path.join(r, f), d) for r, _, fs in os.walk(d) for f in fs
                       if r.startswith(out) or r.startswith(art))
        diag_json = json.load(open(os.path.join(art, "diagnostics.json")))
        cfg_json = json.load(open(os.path.join(art, "config.json")))
    # right singular vectors of the very matrices the pipeline decomposed (same function, same input)
    vh, vh_noise = {}, {}
    combined = quiet(ref_masks.combine_masks, masks, strategy=cfg.svd_mask_strategy, verbose=False) if masks else {}
    for p, b in res["bases"].items():
        cols = []
        for t in tasks:
            dlt = fts[t][p] - base[p]
            m = combined.get(p)
            cols.append(dlt.flatten()[m.flatten()] if m is not None else dlt.flatten())
        T, _ = ref_basis.stack_and_center(cols, cfg.svd_center)
        _, S, Vh = ref_basis.compute_svd(T)
        assert torch.equal(S, b["masked"]["singular_values"])
        vh[p] = Vh
        if b.get("noise") is not None:
            rest = [(fts[t][p] - base[p]).flatten()[~combined[p].flatten()] for t in tasks]
            T, _ = ref_basis.stack_and_center(rest, cfg.svd_center)
            _, S, Vh = ref_basis.compute_svd(T)
            assert torch.equal(S, b["noise"]["singular_values"])
            vh_noise[p] = Vh
    diag = res["diagnostics"]
    return {"tasks": tasks, "base": dict(base), "finetuned": {t: dict(fts[t]) for t in tasks},
            "masks": {t: dict(masks[t]) for t in tasks} if masks is not None else None,
            "config": dict(spec["cfg"]), "performance": synth.performance_table(tasks) if spec.get("perf") else None,
            "merged_state_dict": res["merged_state_dict"], "bases": res["bases"], "compressed": res["compressed"],
            "diagnostics": json.loads(json.dumps(diag, default=lambda o: o.item() if hasattr(o, "item") else list(o))),
            "diagnostics_json": diag_json, "config_json": cfg_json, "files": files, "Vh": vh,
            "Vh_noise": vh_noise, "combined_masks": combined}


def main():
    torch.manual_seed(0)
    torch.save(rtvq_cases(), os.path.join(HERE, "rtvq_golden.pt"))
    torch.save(mask_cases(), os.path.join(HERE, "mask_golden.pt"))
    torch.save(rank_cases(), os.path.join(HERE, "rank_golden.pt"))
    pipe = {name: pipeline_case(name, spec) for name, spec in PIPELINE_CASES.items()}
    torch.save(pipe, os.path.join(HERE, "pipeline_golden.pt"))
    meta = {"torch": torch.__version__, "reference": "mgradyn/SVD-Quantization-Task-Merging @ /root/reference",
            "generator": "tests/golden/make_golden.py", "cases": {k: v["files"] for k, v in pipe.items()}}
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
