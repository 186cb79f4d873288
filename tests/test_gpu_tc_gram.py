"""Tensor-core pass 1 (k9_gram_tc.cu: tcgen05.mma kind::f16, TMEM accumulators) for 16-bit checkpoints against
(a) the CUDA-core pass 1 on the same inputs (SVDQ_TC=0) -- combined masks and counts bit-identical, Grams equal to
fp32 round-off -- and (b) an fp64 Gram of the exactly-rounded task vectors (reference: `ft - base` on 16-bit
tensors, src/svd_hybrid/task_vector_loader.py:142; T^T T of basis.py:63-113,241)."""
import numpy as np
import pytest
import torch

from svd_quantization_task_merging_b200 import synth
from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig

pytestmark = pytest.mark.gpu

SHAPES = {"big": (3, 16384 + 1024 + 40), "chunk": (1024,), "two": (2048,), "odd": (1531,), "tiny": (7,),
          "tile": (16384,), "mid": (130, 257)}


def _job(base, fts, masks, cfg, monkeypatch, tc):
    from svd_quantization_task_merging_b200.engine import MergeJob
    monkeypatch.setenv("SVDQ_TC", "1" if tc else "0")
    return MergeJob(base, fts, masks, cfg, "cuda", diagnostics=False).run()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n_tasks,strategy,mask_p,host_masks", [
    (8, "union", None, False), (8, "intersection", 0.9, False), (5, "majority", 0.5, False), (8, "union", 0.3, True),
    (3, "intersection", 0.8, True), (1, "union", None, False)])
def test_tc_gram_matches_cuda_core_gram(cuda_device, monkeypatch, dtype, n_tasks, strategy, mask_p, host_masks):
    tasks = synth.task_names(n_tasks)
    base, fts = synth.make_checkpoints(SHAPES, tasks, family="parity", seed=3, dtype=dtype, device="cuda")
    if n_tasks >= 3:
        del fts[tasks[1]]["two"]                    # a task that lacks a parameter
    masks = None
    if mask_p is not None:
        masks = synth.make_masks(SHAPES, tasks, mask_p, seed=4, device="cpu" if host_masks else "cuda")
        if n_tasks >= 3:
            del masks[tasks[2]]["mid"]              # a task without a mask for a parameter
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy=strategy, svd_store_artifacts=False,
                          svd_eval_reconstruction=False)
    a = _job(base, fts, masks, cfg, monkeypatch, tc=True)
    b = _job(base, fts, masks, cfg, monkeypatch, tc=False)
    ga, gb = a.groups[dtype], b.groups[dtype]
    assert ga.names == gb.names
    assert torch.equal(ga.t["dm"], gb.t["dm"])
    cma, cmb = a.combined_masks(), b.combined_masks()
    assert cma.keys() == cmb.keys() and all(torch.equal(cma[k], cmb[k]) for k in cma)
    Ga = ga.t["gram_masked"].cpu().numpy().reshape(len(ga.names), n_tasks, n_tasks)
    Gb = gb.t["gram_masked"].cpu().numpy().reshape(len(ga.names), n_tasks, n_tasks)
    worst = 0.0
    for p, name in enumerate(ga.names):
        # fp64 Gram of the exactly-rounded masked task vectors
        cols = []
        for t in tasks:
            if name in fts[t]:
                d = (fts[t][name] - base[name]).double().flatten()          # 16-bit subtract, like the reference
            else:
                d = torch.zeros(base[name].numel(), dtype=torch.float64, device="cuda")
            if name in cma:
                d = d * cma[name].flatten().double()
            cols.append(d)
        T = torch.stack(cols, 1)
        G = (T.T @ T).cpu().numpy()
        scale = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
        ea, eb = np.abs(Ga[p] - G) / scale, np.abs(Gb[p] - G) / scale
        worst = max(worst, ea.max())
        assert ea.max() <= 3e-7, (name, ea.max(), eb.max())
        assert np.abs(Ga[p] - Gb[p]).max() <= 4e-7 * scale.max(), name
    print(f"tensor-core Gram vs fp64: max relative error {worst:.2e}")
    # downstream: same ranks; merged weights agree at the fp32 round-off of the Gram
    fa, fb = a._fetch()[dtype], b._fetch()[dtype]
    assert (fa["info"][:, :4] == fb["info"][:, :4]).all()
    ma, mb = a.merged_state_dict(), b.merged_state_dict()
    for k in ma:
        fin = torch.isfinite(mb[k].float())
        assert torch.equal(torch.isfinite(ma[k].float()), fin)
        if fin.any() and k not in ("tiny",):
            da = (ma[k].float() - base[k].float())[fin].double()
            db = (mb[k].float() - base[k].float())[fin].double()
            assert (da - db).norm() <= 2e-3 * db.norm() + 1e-12, k


def test_tc_gram_is_deterministic_and_placement_independent(cuda_device, monkeypatch):
    """Per-tile partials are a fixed function of the tile: two runs, and a run on a sub-set of the parameters,
    give identical bits (what the sharded path relies on)."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    monkeypatch.setenv("SVDQ_TC", "1")
    tasks = synth.task_names(8)
    base, fts = synth.make_checkpoints(SHAPES, tasks, family="parity", seed=5, dtype=torch.bfloat16, device="cuda")
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_store_artifacts=False, svd_eval_reconstruction=False)
    a = MergeJob(base, fts, None, cfg, "cuda", diagnostics=False).run()
    b = MergeJob(base, fts, None, cfg, "cuda", diagnostics=False).run()
    c = MergeJob(base, fts, None, cfg, "cuda", diagnostics=False, param_filter=["big", "odd"]).run()
    ga, gb, gc = (j.groups[torch.bfloat16] for j in (a, b, c))
    assert torch.equal(ga.t["gram_masked"], gb.t["gram_masked"])
    for name in gc.names:
        assert torch.equal(gc.t["gram_masked"][gc.names.index(name)], ga.t["gram_masked"][ga.names.index(name)])
    ma, mc = a.merged_state_dict(), c.merged_state_dict()
    assert all(torch.equal(ma[k], mc[k]) for k in mc)


@pytest.mark.parametrize("fp16_bases", [True, False])
@pytest.mark.parametrize("n_tasks,strategy,mask_p,weighting", [
    (8, "union", None, "uniform"), (8, "intersection", 0.9, "performance"), (5, "majority", 0.5, "uniform"),
    (3, "union", 0.5, "uniform"), (1, "union", None, "uniform")])
def test_tc_merge_matches_cuda_core_merge(cuda_device, monkeypatch, fp16_bases, n_tasks, strategy, mask_p, weighting):
    """Tensor-core pass 2 (k10_merge_tc.cu, SVDQ_TC bit 1) against the CUDA-core pass 2 on identical coefficients
    (pass 1 on CUDA cores in both runs): same ranks / codes by construction, merged weights equal to fp32 round-off
    of the basis rows (an fp16 rounding of a basis entry may land on the other side: 5e-4 relative on that entry),
    untouched elements bit-identical."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    tasks = synth.task_names(n_tasks)
    dtype = torch.bfloat16
    base, fts = synth.make_checkpoints(SHAPES, tasks, family="parity", seed=13, dtype=dtype, device="cuda")
    if n_tasks >= 3:
        del fts[tasks[1]]["two"]
    masks = synth.make_masks(SHAPES, tasks, mask_p, seed=14, device="cuda") if mask_p is not None else None
    perf = synth.performance_table(tasks) if weighting == "performance" else None
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy=strategy, svd_weighting=weighting,
                          svd_fp16=fp16_bases, svd_store_artifacts=False, svd_eval_reconstruction=False)
    out = {}
    for mode in ("2", "0"):
        monkeypatch.setenv("SVDQ_TC", mode)
        job = MergeJob(base, fts, masks, cfg, "cuda", diagnostics=False, performance=perf).run()
        out[mode] = (job, job.merged_state_dict())
    (ja, ma), (jb, mb) = out["2"], out["0"]
    fa, fb = ja._fetch()[dtype], jb._fetch()[dtype]
    for key in ("info", "chigh", "codes", "cbar"):
        assert np.array_equal(fa[key], fb[key], equal_nan=True), key
    cm = ja.combined_masks()
    worst = 0.0
    for k in mb:
        a32, b32, base32 = ma[k].float(), mb[k].float(), base[k].float()
        fin = torch.isfinite(b32)
        assert torch.equal(torch.isfinite(a32), fin), k
        if k in cm:
            assert torch.equal(a32[~cm[k]], base32[~cm[k]]), k            # outside the mask: merged == base exactly
        da, db = (a32 - base32)[fin].double(), (b32 - base32)[fin].double()
        if db.numel() and db.norm() > 0:
            err = float((da - db).norm() / db.norm())
            worst = max(worst, err)
            assert err <= (2e-5 if fp16_bases else 2e-6), (k, err)
    print(f"tensor-core pass 2 vs CUDA-core pass 2: max rel L2 of the merged delta {worst:.2e}")
