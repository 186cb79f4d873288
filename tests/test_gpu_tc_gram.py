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


WIDE_SHAPES = {"big": (3, 12288 + 1024 + 40), "chunk": (1024,), "odd": (1531,), "tiny": (7,), "tile": (12288,),
               "two_tiles": (2 * 12288 + 128,), "mid": (130, 257), "stage": (512,), "sub": (128,)}


@pytest.mark.parametrize("n_tasks,strategy,mask_p,weighting,noise", [
    (20, "union", 0.5, "uniform", False), (17, "intersection", 0.97, "cluster", False),
    (21, "majority", 0.5, "uniform", True), (18, "union", None, "uniform", False),
    (19, "union", 0.3, "cluster", False), (20, "union", None, "cluster", False),
    (24, "majority", 0.5, "uniform", False)])        # above 21 tasks: the CUDA-core kernel serves both runs
def test_wide_tc_gram_matches_fp64_and_cuda_core(cuda_device, monkeypatch, n_tasks, strategy, mask_p, weighting, noise):
    """Tensor-core single-pass Gram of the wide path (k12_gram_wide_tc.cu, SVDQ_TC bit 2: fp32 task vectors as three
    exact bf16 pieces on tcgen05) against an fp64 Gram of the masked task vectors and against the CUDA-core kernel
    (k8_gram_staged.cu) on the same inputs: same accuracy class, same ranks, merged weights equal at the round-off
    of the Gram.  Covers all three mask modes (rows inside the mask, all rows for cluster weighting, rows outside
    for the noise region), parameters that end inside a stage / a tile buffer, and tasks that lack a parameter."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    tasks = synth.task_names(n_tasks)
    base, fts = synth.make_checkpoints(WIDE_SHAPES, tasks, family="parity", seed=23, device="cuda")
    del fts[tasks[1]]["odd"]
    masks = synth.make_masks(WIDE_SHAPES, tasks, mask_p, seed=24, device="cuda") if mask_p is not None else None
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy=strategy or "union",
                          svd_weighting=weighting, svd_include_noise=noise, svd_store_artifacts=False,
                          svd_eval_reconstruction=False)
    out = {}
    for mode in ("4", "0"):
        monkeypatch.setenv("SVDQ_TC", mode)
        out[mode] = MergeJob(base, fts, masks, cfg, "cuda", diagnostics=False).run()
    a, b = out["4"], out["0"]
    ga, gb = a.groups[torch.float32], b.groups[torch.float32]
    cm = a.combined_masks()
    worst_a = worst_b = 0.0
    keys = ["gram_masked"] + (["gram_all"] if weighting == "cluster" else [])
    for key in keys:
        Ga = ga.t[key].cpu().numpy().reshape(-1, n_tasks, n_tasks)
        Gb = gb.t[key].cpu().numpy().reshape(-1, n_tasks, n_tasks)
        if key == "gram_all":                               # one whole-model Gram
            Ga, Gb = Ga.sum(0, keepdims=True), Gb.sum(0, keepdims=True)
        acc = np.zeros((n_tasks, n_tasks))
        for p, name in enumerate(ga.names):
            cols = []
            for t in tasks:
                d = (fts[t][name] - base[name]).double().flatten() if name in fts[t] else \
                    torch.zeros(base[name].numel(), dtype=torch.float64, device="cuda")
                if key == "gram_masked" and name in cm:
                    d = d * cm[name].flatten().double()
                cols.append(d)
            T = torch.stack(cols, 1)
            G = (T.T @ T).cpu().numpy()
            if key == "gram_all":
                acc += G
                continue
            scale = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
            ea, eb = (np.abs(Ga[p] - G) / scale).max(), (np.abs(Gb[p] - G) / scale).max()
            worst_a, worst_b = max(worst_a, ea), max(worst_b, eb)
            assert ea <= 4e-7, (key, name, ea, eb)
        if key == "gram_all":
            scale = np.sqrt(np.outer(np.diag(acc), np.diag(acc)))
            assert (np.abs(Ga[0] - acc) / scale).max() <= 4e-7
    print(f"wide Gram vs fp64, max relative error: tensor cores {worst_a:.2e}, CUDA cores {worst_b:.2e}")
    fa, fb = a._fetch()[torch.float32], b._fetch()[torch.float32]
    assert (fa["info"][:, :4] == fb["info"][:, :4]).all()
    ma, mb = a.merged_state_dict(), b.merged_state_dict()
    for k in ma:
        fin = torch.isfinite(mb[k])
        assert torch.equal(torch.isfinite(ma[k]), fin), k
        if fin.any() and k != "tiny":
            da, db = (ma[k] - base[k])[fin].double(), (mb[k] - base[k])[fin].double()
            assert (da - db).norm() <= 2e-3 * db.norm() + 1e-12, k


def test_wide_tc_gram_is_deterministic_and_placement_independent(cuda_device, monkeypatch):
    from svd_quantization_task_merging_b200.engine import MergeJob
    monkeypatch.setenv("SVDQ_TC", "4")
    tasks = synth.task_names(20)
    base, fts = synth.make_checkpoints(WIDE_SHAPES, tasks, family="parity", seed=5, device="cuda")
    masks = synth.make_masks(WIDE_SHAPES, tasks, 0.5, seed=6, device="cuda")
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_store_artifacts=False, svd_eval_reconstruction=False)
    a = MergeJob(base, fts, masks, cfg, "cuda", diagnostics=False).run()
    b = MergeJob(base, fts, masks, cfg, "cuda", diagnostics=False).run()
    c = MergeJob(base, fts, masks, cfg, "cuda", diagnostics=False, param_filter=["big", "odd", "two_tiles"]).run()
    ga, gb, gc = (j.groups[torch.float32] for j in (a, b, c))
    assert torch.equal(ga.t["gram_masked"], gb.t["gram_masked"])
    for name in gc.names:
        assert torch.equal(gc.t["gram_masked"][gc.names.index(name)], ga.t["gram_masked"][ga.names.index(name)])
    ma, mc = a.merged_state_dict(), c.merged_state_dict()
    assert all(torch.equal(ma[k], mc[k]) for k in mc)


@pytest.mark.parametrize("n_tasks,strategy,mask_p,weighting,center", [
    (20, "union", 0.5, "uniform", True), (17, "intersection", 0.97, "cluster", True),
    (21, "majority", 0.5, "performance", True), (18, "union", None, "uniform", True),
    (19, "union", 0.3, "uniform", False)])
def test_wide_tc_merge_matches_cuda_core_merge(cuda_device, monkeypatch, n_tasks, strategy, mask_p, weighting, center):
    """Tensor-core pass 2 of the wide path (k13_merge_wide_tc.cu, SVDQ_TC bit 3: tau as three exact bf16 pieces times
    the centred projection matrix as three bf16 pieces on tcgen05, TMEM epilogue) against the CUDA-core pass 2
    (k6_reconstruct_merge) on identical coefficients (pass 1 on CUDA cores in both runs): same ranks / codes by
    construction, merged weights equal to fp32 round-off of the basis rows (an fp16 rounding of a basis entry may land
    on the other side: 5e-4 relative on that entry), untouched elements bit-identical.  Covers parameters that end
    inside a stage / a tile buffer, a task that lacks a parameter, and parameters without a basis."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    tasks = synth.task_names(n_tasks)
    base, fts = synth.make_checkpoints(WIDE_SHAPES, tasks, family="parity", seed=33, device="cuda")
    del fts[tasks[1]]["odd"]
    masks = synth.make_masks(WIDE_SHAPES, tasks, mask_p, seed=34, device="cuda") if mask_p is not None else None
    perf = synth.performance_table(tasks) if weighting == "performance" else None
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy=strategy, svd_weighting=weighting,
                          svd_center=center, svd_fp16=True, svd_store_artifacts=False, svd_eval_reconstruction=False)
    out = {}
    for mode in ("8", "0"):
        monkeypatch.setenv("SVDQ_TC", mode)
        job = MergeJob(base, fts, masks, cfg, "cuda", diagnostics=False, performance=perf).run()
        out[mode] = (job, job.merged_state_dict())
    (ja, ma), (jb, mb) = out["8"], out["0"]
    fa, fb = ja._fetch()[torch.float32], jb._fetch()[torch.float32]
    for key in ("info", "chigh", "codes", "cbar"):
        assert np.array_equal(fa[key], fb[key], equal_nan=True), key
    cm = ja.combined_masks()
    worst = 0.0
    for k in mb:
        a32, b32 = ma[k], mb[k]
        fin = torch.isfinite(b32)
        assert torch.equal(torch.isfinite(a32), fin), k
        if k in cm:
            assert torch.equal(a32[~cm[k]], base[k][~cm[k]]), k            # outside the mask: merged == base exactly
        da, db = (a32 - base[k])[fin].double(), (b32 - base[k])[fin].double()
        if db.numel() and db.norm() > 0:
            err = float((da - db).norm() / db.norm())
            worst = max(worst, err)
            assert err <= 2e-5, (k, err)
    print(f"wide tensor-core pass 2 vs CUDA-core pass 2: max rel L2 of the merged delta {worst:.2e}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n_tasks,strategy,mask_p", [(8, "intersection", 0.9), (5, "majority", 0.5), (8, "union", None), (3, "intersection", 0.6)])
def test_compacting_diag_pass_matches_plain_walk(cuda_device, monkeypatch, dtype, n_tasks, strategy, mask_p):
    """Pass 2 with fused diagnostics, compacting kernel (k3c_merge_diag_compact.cu: only the elements inside the combined
    mask go through the arithmetic) against the non-compacting kernel on identical coefficients: merged weights
    bit-identical (same per-element instruction sequence), per-task error figures equal to the fp32 round-off of a
    different summation order.  SVDQ_COMPACT_DIAG = 2 forces compaction for every tile, 1 chooses per tile, 0 never."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    tasks = synth.task_names(n_tasks)
    base, fts = synth.make_checkpoints(SHAPES, tasks, family="parity", seed=43, dtype=dtype, device="cuda")
    if n_tasks >= 3:
        del fts[tasks[1]]["two"]
    masks = synth.make_masks(SHAPES, tasks, mask_p, seed=44, device="cuda") if mask_p is not None else None
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy=strategy, svd_store_artifacts=False,
                          svd_eval_reconstruction=True)
    out = {}
    for mode in ("2", "1", "0"):
        monkeypatch.setenv("SVDQ_COMPACT_DIAG", mode)
        job = MergeJob(base, fts, masks, cfg, "cuda", diagnostics=True).run()
        res = job.results()
        out[mode] = (job.merged_state_dict(), res["diagnostics"]["per_parameter"])
    for mode in ("2", "1"):
        ma, da = out[mode]
        mb, db = out["0"]
        for k in mb:
            assert torch.equal(ma[k].view(torch.int32 if ma[k].element_size() == 4 else torch.int16),
                               mb[k].view(torch.int32 if mb[k].element_size() == 4 else torch.int16)), (mode, k)
        for name, pr in db.items():
            for task, er in pr["reconstruction_errors"].items():
                for key, v in er.items():
                    w = da[name]["reconstruction_errors"][task][key]
                    if v != v:
                        assert w != w, (mode, name, task, key)
                    else:
                        assert abs(w - v) <= 2e-5 * abs(v) + 1e-12, (mode, name, task, key, w, v)


@pytest.mark.parametrize("dtype,fp16_bases", [(torch.float32, True), (torch.float32, False), (torch.bfloat16, True)])
@pytest.mark.parametrize("n_tasks,strategy,mask_p,center", [(8, "intersection", 0.9, True), (5, "majority", 0.5, True),
                                                            (8, "union", None, True), (3, "union", 0.3, False)])
def test_fused_basis_write_out_equals_k5(cuda_device, monkeypatch, dtype, fp16_bases, n_tasks, strategy, mask_p, center):
    """The reference's default settings (svd_eval_reconstruction AND svd_store_artifacts): pass 2 writes the artifact
    bases itself (svdq_reconstruct_merge_basis: U_high / U_low / mean compacted to the masked rows) instead of a third
    pass over the inputs (K5, svdq_write_basis).  Same bits in the bases, the merged model and the coefficients;
    diagnostics equal to the round-off of the summation order."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    tasks = synth.task_names(n_tasks)
    base, fts = synth.make_checkpoints(SHAPES, tasks, family="parity", seed=53, dtype=dtype, device="cuda")
    if n_tasks >= 3:
        del fts[tasks[1]]["two"]
    masks = synth.make_masks(SHAPES, tasks, mask_p, seed=54, device="cuda") if mask_p is not None else None
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy=strategy, svd_center=center,
                          svd_fp16=fp16_bases, svd_store_artifacts=True, svd_eval_reconstruction=True)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SVDQ_FUSED_BASIS", mode)
        job = MergeJob(base, fts, masks, cfg, "cuda", diagnostics=True, materialize_bases=True).run()
        assert (job._fused_basis_buffers() is not None) == (mode == "1")
        res = job.results()
        out[mode] = (job, res)
    (ja, ra), (jb, rb) = out["1"], out["0"]
    n_checked = 0
    for name in rb["bases"]:
        ba, bb = ra["bases"][name]["masked"], rb["bases"][name]["masked"]
        assert ba["k"] == bb["k"] and ba["D"] == bb["D"]
        for key in ("U_high", "U_low"):
            assert ba[key].shape == bb[key].shape and ba[key].dtype == bb[key].dtype, (name, key)
            bits = torch.int16 if ba[key].element_size() == 2 else torch.int32
            assert torch.equal(ba[key].contiguous().view(bits), bb[key].contiguous().view(bits)), (name, key)
        if bb.get("mean") is not None:
            assert torch.equal(ba["mean"].view(torch.int32), bb["mean"].view(torch.int32)), name
        else:
            assert ba.get("mean") is None
        n_checked += 1
    assert n_checked >= 4
    for k in rb["merged_state_dict"]:
        a_, b_ = ra["merged_state_dict"][k], rb["merged_state_dict"][k]
        bits = torch.int16 if a_.element_size() == 2 else torch.int32
        assert torch.equal(a_.view(bits), b_.view(bits)), k
