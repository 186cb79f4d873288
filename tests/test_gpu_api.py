"""GPU tests of the reference-facing surface: the CUDA path against golden outputs of the REAL
reference, the fine-grained operator mirrors, bases as subspaces, the on-disk pipeline + artifact
layout + reload, and world-size invariance of the parameter-sharded path."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import svd_hybrid_ref as R
from svd_quantization_task_merging_b200 import sharding, synth
from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig
from tests import parity

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    f = "pipeline_golden_wide.pt" if name.startswith("wide") else "pipeline_golden.pt"     # > 16 tasks: own file
    return torch.load(os.path.join(GOLD, f), weights_only=False)[name]


# ---- the CUDA path against the real reference's outputs -------------------------------------------------
@pytest.mark.parametrize("name", ["union_uniform", "majority_performance_3stage", "intersection_cluster",
                                  "nomask_fp32_nocenter", "majority_noise_uniform", "union_noise_cluster_3stage",
                                  "wide20_union_uniform", "wide24_majority_cluster"])
def test_cuda_path_against_reference_golden(cuda_device, name):
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    case = _gold(name)
    cfg = SVDHybridConfig(tasks=case["tasks"], svd_max_rank=64, svd_store_artifacts=False, **case["config"])
    ref = parity.golden_as_reference(case)
    res = merge_state_dicts(case["base"], case["finetuned"], case["masks"], cfg, "cuda", sign_ref=case["Vh"],
                            sign_ref_noise=case.get("Vh_noise"), performance=case["performance"],
                            cluster_assignments=ref["cluster_assignments"])
    rep = parity.compare_run(ref, res)
    print(name, rep)
    assert res["diagnostics"]["task_weights"] == ref["weights"]


def test_cuda_path_reproduces_reference_nans_golden(cuda_device):
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    case = _gold("iid_degenerate_nan")
    cfg = SVDHybridConfig(tasks=case["tasks"], svd_max_rank=64, svd_store_artifacts=False, **case["config"])
    res = merge_state_dicts(case["base"], case["finetuned"], None, cfg, "cuda")
    for p, m in case["merged_state_dict"].items():
        assert parity.nan_positions_equal(res["merged_state_dict"][p], m), p
        assert torch.isnan(m).all()
        assert res["bases"].meta(p)["k"] == case["bases"][p]["masked"]["k"] == 7


def test_cluster_partition_golden_from_gram(cuda_device):
    """k-means on the K1 whole-model Gram reproduces the reference's own cluster partition."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    case = _gold("intersection_cluster")
    cfg = SVDHybridConfig(tasks=case["tasks"], svd_max_rank=64, svd_store_artifacts=False, **case["config"])
    gold = case["diagnostics"]["cluster_assignments"]
    # default backend = the reference's KMeans(k, random_state=42, n_init=10) procedure restated in libsvdq
    # (svdq_host_kmeans) on an isometric embedding of the Gram; these toy tasks are unclustered, so only the
    # reference's own procedure (local optimum included) reproduces its partition
    job = MergeJob(case["base"], case["finetuned"], case["masks"], cfg, "cuda").run()
    mine, ts = job.cluster_assignments, case["tasks"]
    assert all((mine[a] == mine[b]) == (gold[a] == gold[b]) for a in ts for b in ts)
    assert mine == gold                                  # same label numbering as well


# ---- bases as subspaces ------------------------------------------------------------------------------------
@pytest.mark.parametrize("fp16", [False, True])
def test_bases_match_as_subspaces(cuda_device, fp16):
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    tasks = synth.task_names(8)
    base, fts = synth.make_checkpoints(parity.MEDIUM_SHAPES, tasks, family="parity", seed=11)
    masks = synth.make_masks(parity.MEDIUM_SHAPES, tasks, 0.5, seed=12)
    ref_cfg, cfg = parity.make_cfgs(tasks, svd_energy_threshold=0.9, svd_fp16=fp16)
    ref = R.run_reference_path(base, fts, masks, ref_cfg)
    res = merge_state_dicts(base, fts, masks, cfg, "cuda", sign_ref={p: b["Vh"] for p, b in ref["bases"].items()},
                            materialize_bases=True)
    tol = 2e-3 if fp16 else parity.TOL_ANGLE
    for p, rb in ref["bases"].items():
        nb = res["bases"][p]["masked"]
        assert list(nb.keys()) == ["U_high", "U_low", "singular_values", "k", "mean", "energy_retained", "D", "N"]
        assert nb["U_high"].shape == rb["U_high"].shape and nb["U_low"].shape == rb["U_low"].shape
        assert nb["U_high"].dtype == rb["U_high"].dtype == (torch.float16 if fp16 else torch.float32)
        assert nb["mean"].shape == rb["mean"].shape and nb["singular_values"].shape == rb["singular_values"].shape
        assert torch.allclose(nb["mean"].cpu(), rb["mean"], rtol=1e-5, atol=1e-9)
        k = rb["k"]
        assert parity.max_principal_sine(nb["U_high"].float(), rb["U_high"].float()) <= tol, p
        # all numerically non-null directions (the centred task matrix has rank N-1: the last left vector
        # of LAPACK is round-off noise, here it is a zero column)
        rank = int((rb["singular_values"] > 1e-5 * rb["singular_values"][0]).sum())
        U_ref = torch.cat([rb["U_high"], rb["U_low"]], 1).float()[:, :rank]
        U_new = torch.cat([nb["U_high"], nb["U_low"]], 1).float()[:, :rank]
        assert parity.max_principal_sine(U_new, U_ref) <= tol, p
        # sign-aligned columns agree entry-wise
        # LAPACK's own fp32 error on the weakest direction is ~eps * sigma_1 / sigma_j ~ 2e-6 of |u|
        assert (U_new.cpu() - U_ref).abs().max() <= (2e-3 if fp16 else 1e-4) * U_ref.abs().max()
        # orthonormality of what we wrote
        eye = U_new.double().T @ U_new.double()
        assert (eye.cpu() - torch.eye(rank, dtype=torch.float64)).abs().max() < (2e-3 if fp16 else 1e-5)


# ---- fine-grained operator mirrors -----------------------------------------------------------------------------
def test_mask_strategies_truth_tables_and_golden(cuda_device):
    from src.svd_hybrid.mask_loader import (combine_masks, compute_intersection_mask, compute_majority_mask,
                                            compute_union_mask)
    a = torch.tensor([True, True, False, False])
    b = torch.tensor([True, False, True, False])
    c = torch.tensor([True, False, False, False])
    assert compute_union_mask([a, b]).tolist() == [True, True, True, False]
    assert compute_intersection_mask([a, b]).tolist() == [True, False, False, False]
    assert compute_majority_mask([a, b, c]).tolist() == [True, False, False, False]
    assert compute_majority_mask([a, b]).tolist() == [True, True, True, False]          # votes >= 0.5 * N
    assert compute_union_mask([a]).tolist() == a.tolist() and compute_union_mask([a]).dtype == torch.bool
    for fn in (compute_union_mask, compute_intersection_mask, compute_majority_mask):
        with pytest.raises(ValueError):
            fn([])
    assert combine_masks({}, "union", verbose=False) == {}
    with pytest.raises(ValueError):
        combine_masks({"t": {"w": a}}, "xor", verbose=False)
    for case in torch.load(os.path.join(GOLD, "mask_golden.pt"), weights_only=False):
        for strat in ("union", "intersection", "majority"):
            mine = combine_masks(case["masks"], strategy=strat, verbose=False)
            assert sorted(mine) == sorted(case[strat])
            for name, m in case[strat].items():
                assert mine[name].dtype == torch.bool and mine[name].shape == m.shape and torch.equal(mine[name], m)
    big = [torch.rand(3, 100003, generator=torch.Generator().manual_seed(i)) < 0.5 for i in range(5)]
    for strat in ("union", "intersection", "majority"):
        assert torch.equal(combine_masks({str(i): {"w": m} for i, m in enumerate(big)}, strat, verbose=False)["w"],
                           R.combine_mask_list(big, strat))


def test_construct_basis_compress_reconstruct_with_mean(cuda_device):
    """tests/test_mean_handling.py of the reference: centred basis -> compress -> dequantise -> reconstruct must
    add the mean back; uncentred path < 5 %; exact algebra U_h c_h + U_l c_l (+ mean)."""
    from src.svd_hybrid.basis import compute_svd, construct_basis
    from src.svd_hybrid.compress import compress_single_task, project_to_basis
    from src.svd_hybrid.merge import reconstruct_from_coefficients
    from src.svd_hybrid.rtvq import RTVQQuantizer
    g = torch.Generator().manual_seed(0)
    D, N = 600, 8
    q, _ = torch.linalg.qr(torch.randn(N, N, generator=g))
    deltas = list(((q * (0.6 ** torch.arange(N))) @ torch.randn(N, D, generator=g) + 1.0))      # common mean
    quant = RTVQQuantizer(num_bits=8, num_stages=2)
    b = construct_basis(deltas, energy_threshold=0.9, max_rank=None, center=True, device="cpu", verbose=False)
    assert b["mean"].shape == (D, 1) and b["U_high"].shape == (D, b["k"]) and b["U_low"].shape == (D, N - b["k"])
    assert b["D"] == D and b["N"] == N and b["U_high"].device.type == "cpu"
    ob = R.build_basis(deltas, 0.9, None, True)
    assert b["k"] == ob["k"]
    assert (b["singular_values"] - ob["singular_values"]).abs().max() <= 5 * parity.TOL_S * ob["singular_values"][0]
    for d in deltas[:3]:
        art = compress_single_task(d, b["U_high"], b["U_low"], quant, "cpu", mean=b["mean"])
        assert art["c_high_fp16"].dtype == torch.float16 and art["c_high_fp16"].device.type == "cpu"
        c_lo = quant.dequantize(art["c_low_quant"])
        with_mean = reconstruct_from_coefficients(art["c_high_fp16"].float(), c_lo, b["U_high"], b["U_low"], "cpu",
                                                  mean=b["mean"])
        without = reconstruct_from_coefficients(art["c_high_fp16"].float(), c_lo, b["U_high"], b["U_low"], "cpu")
        e1 = ((with_mean - d).norm() / d.norm()).item()
        e0 = ((without - d).norm() / d.norm()).item()
        assert e1 < 0.01 and e0 > 10 * e1
    # uncentred path
    b2 = construct_basis(deltas, energy_threshold=0.99, center=False, device="cpu", verbose=False)
    assert b2["mean"] is None
    art = compress_single_task(deltas[0], b2["U_high"], b2["U_low"], quant, "cpu", mean=None)
    rec = reconstruct_from_coefficients(art["c_high_fp16"].float(), quant.dequantize(art["c_low_quant"]),
                                        b2["U_high"], b2["U_low"], "cpu")
    assert ((rec - deltas[0]).norm() / deltas[0].norm()).item() < 0.05
    # exact algebra on a full orthogonal basis (tests/test_mean_handling.py:206-312)
    Q, _ = torch.linalg.qr(torch.randn(50, 50, generator=g))
    x = torch.randn(50, generator=g)
    ch, cl = project_to_basis(x, Q[:, :10], Q[:, 10:])
    assert torch.allclose(reconstruct_from_coefficients(ch, cl, Q[:, :10], Q[:, 10:], "cpu"), x, atol=1e-5)
    m = torch.randn(50, 1, generator=g)
    assert torch.allclose(reconstruct_from_coefficients(ch, cl, Q[:, :10], Q[:, 10:], "cpu", mean=m), x + m.squeeze(), atol=1e-5)
    # compute_svd keeps the device and reproduces the matrix
    M = torch.randn(300, 6, generator=g)
    U, S, Vh = compute_svd(M)
    assert U.device == M.device and U.shape == (300, 6) and S.shape == (6,) and Vh.shape == (6, 6)
    assert torch.allclose((U * S) @ Vh, M, atol=1e-4)
    assert torch.allclose(S, torch.linalg.svdvals(M), rtol=1e-5)
    Uc, Sc, Vc = compute_svd(M.cuda())
    assert Uc.is_cuda and Sc.is_cuda and Vc.is_cuda


# ---- the on-disk pipeline -------------------------------------------------------------------------------------
def _write_case(tmp, case):
    ck, md = tmp / "ckpt", tmp / "masks"
    ck.mkdir()
    md.mkdir()
    torch.save(dict(case["base"]), tmp / "base.pt")
    for t in case["tasks"]:
        torch.save(dict(case["finetuned"][t]), ck / f"{t}.pt")
        if case["masks"] is not None:
            torch.save(dict(case["masks"][t]), md / f"{t}_mask.pt")
    return ck, md


@pytest.mark.parametrize("name", ["union_uniform", "intersection_cluster", "majority_noise_uniform", "wide20_union_uniform"])
def test_pipeline_on_disk_writes_reference_layout(cuda_device, tmp_path, name, monkeypatch):
    from src.svd_hybrid.cli import run_svd_hybrid_pipeline
    from src.svd_hybrid.storage import load_all_artifacts
    case = _gold(name)
    ck, md = _write_case(tmp_path, case)
    cfg = SVDHybridConfig(tasks=case["tasks"], checkpoint_dir=str(ck), base_model_path=str(tmp_path / "base.pt"),
                          mask_dir=str(md), svd_store_artifacts=True, svd_eval_reconstruction=True, svd_max_rank=64,
                          output_dir=str(tmp_path / "out"), artifact_dir=str(tmp_path / "art"), device="cuda",
                          **case["config"])
    res = run_svd_hybrid_pipeline(cfg, verbose=False)
    assert list(res.keys()) == ["merged_state_dict", "diagnostics", "bases", "compressed"]
    files = sorted(os.path.relpath(os.path.join(r, f), tmp_path) for r, _, fs in os.walk(tmp_path) for f in fs
                   if os.path.relpath(r, tmp_path).split(os.sep)[0] in ("out", "art"))
    extra = {os.path.join("art", "combined_masks.pt")} if case["masks"] is not None else set()
    assert set(files) == set(case["files"]) | extra     # the files the reference wrote (+ the additive mask file)
    art = load_all_artifacts(str(tmp_path / "art"))
    gold_diag = case["diagnostics_json"]
    assert set(art["diagnostics"].keys()) == set(gold_diag.keys())
    assert json.load(open(tmp_path / "art" / "config.json")).keys() == case["config_json"].keys()
    for p, gb in case["bases"].items():
        nb = art["bases"][p]["masked"]
        g = gb["masked"]
        assert list(nb.keys()) == list(g.keys())
        for key in ("U_high", "U_low", "singular_values", "mean"):
            assert nb[key].shape == g[key].shape and nb[key].dtype == g[key].dtype and nb[key].device.type == "cpu"
        assert nb["k"] == g["k"] and nb["D"] == g["D"] and nb["N"] == g["N"]
        # noise region artifacts (storage.py:92-104,166-170): present exactly where the reference wrote them
        assert ("noise" in art["bases"][p]) == (gb.get("noise") is not None), p
        if gb.get("noise") is not None:
            nn, gn = art["bases"][p]["noise"], gb["noise"]
            assert list(nn.keys()) == list(gn.keys())
            for key in ("U_high", "U_low", "singular_values", "mean"):
                assert nn[key].shape == gn[key].shape and nn[key].dtype == gn[key].dtype
            assert nn["k"] == gn["k"] and nn["D"] == gn["D"] and nn["N"] == gn["N"]
        for t in case["tasks"]:
            gu = case["compressed"][p][t].get("unmasked")
            assert ("unmasked" in art["compressed"][p][t]) == (gu is not None)
            if gu is not None:
                au = art["compressed"][p][t]["unmasked"]
                assert au["c_high_fp16"].shape == gu["c_high_fp16"].shape
                assert au["c_low_quant"]["original_shape"] == gu["c_low_quant"]["original_shape"]
            a, b = art["compressed"][p][t]["masked"], case["compressed"][p][t]["masked"]
            assert a["c_high_fp16"].dtype == torch.float16 and a["c_high_fp16"].shape == b["c_high_fp16"].shape
            qa, qb = a["c_low_quant"], b["c_low_quant"]
            assert list(qa.keys()) == list(qb.keys()) and qa["num_bits"] == qb["num_bits"]
            assert qa["original_shape"] == qb["original_shape"] and qa["original_dtype"] == qb["original_dtype"]
            for x, y in zip(qa["payloads"], qb["payloads"]):
                assert list(x.keys()) == list(y.keys()) and x["quantized"].dtype == torch.uint8
                assert x["quantized"].shape == y["quantized"].shape and x["scale"].ndim == 0 and x["zero_point"].ndim == 0
    merged = torch.load(tmp_path / "out" / "merged_state_dict.pt", weights_only=False)
    for p, m in case["merged_state_dict"].items():
        assert merged[p].shape == m.shape and merged[p].dtype == m.dtype
        d_ref = m - case["base"][p]
        assert parity.rel_l2(merged[p] - case["base"][p], d_ref) < 5e-2     # no sign hint: RTVQ-noise level
    assert json.load(open(tmp_path / "out" / "weights.json")) == case["diagnostics"]["task_weights"]


def test_pipeline_on_disk_20_tasks_with_artifacts(cuda_device, tmp_path):
    """The reference's default settings (store artifacts + diagnostics) on a 20-task merge: the wide path writes
    bases / coefficients / diagnostics / merged model in the same layout, and the reload path re-merges them."""
    from src.svd_hybrid.cli import run_svd_hybrid_pipeline
    from src.svd_hybrid.reload import reconstruct_from_artifacts
    from svd_quantization_task_merging_b200 import synth
    shapes = {"a.weight": (96, 40), "a.bias": (1500,), "ln.weight": (257,)}
    tasks = synth.task_names(20)
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=11)
    masks = synth.make_masks(shapes, tasks, 0.5, seed=12)
    case = {"base": base, "finetuned": fts, "masks": masks, "tasks": tasks}
    ck, md = _write_case(tmp_path, case)
    cfg = SVDHybridConfig(tasks=tasks, checkpoint_dir=str(ck), base_model_path=str(tmp_path / "base.pt"),
                          mask_dir=str(md), svd_mask_strategy="majority", svd_store_artifacts=True,
                          svd_eval_reconstruction=True, output_dir=str(tmp_path / "out"),
                          artifact_dir=str(tmp_path / "art"), device="cuda")
    res = run_svd_hybrid_pipeline(cfg, verbose=False)
    for p in shapes:
        assert os.path.exists(tmp_path / "art" / "basis" / f"{p}.pt")
        assert os.path.exists(tmp_path / "art" / "coeffs" / f"{p}.pt")
        b = torch.load(tmp_path / "art" / "basis" / f"{p}.pt", weights_only=False)["masked"]
        assert b["U_high"].shape == (b["D"], b["k"]) and b["U_low"].shape[1] == b["singular_values"].numel() - b["k"]
        assert b["N"] == 20 and b["U_high"].dtype == torch.float16
    assert set(res["diagnostics"]["per_parameter"]) == set(shapes)
    # oracle on the same inputs: merged weights at RTVQ-noise level (no sign hint through the file interface)
    ref = R.run_reference_path(base, fts, masks, R.RefConfig(tasks=tasks, svd_mask_strategy="majority"))
    for p in shapes:
        d_ref = ref["merged_state_dict"][p] - base[p]
        assert parity.rel_l2(res["merged_state_dict"][p].cpu() - base[p], d_ref) < 5e-2
    # reload re-merges the masked run from the artifacts (uses the additive combined_masks.pt)
    out = reconstruct_from_artifacts(str(tmp_path / "art"), str(tmp_path / "base.pt"), str(tmp_path / "re.pt"), "cpu")
    assert set(out["merged_state_dict"]) == set(base)
    for p in shapes:
        assert torch.allclose(out["merged_state_dict"][p], res["merged_state_dict"][p].cpu(), rtol=1e-4, atol=1e-6), p


def test_reload_from_artifacts_unmasked(cuda_device, tmp_path):
    """reload.reconstruct_from_artifacts re-merges from stored bases + codes (reference reload.py:142-238)."""
    from src.svd_hybrid.cli import run_svd_hybrid_pipeline
    from src.svd_hybrid.reload import reconstruct_from_artifacts
    case = _gold("nomask_fp32_nocenter")
    ck, _ = _write_case(tmp_path, case)
    cfg = SVDHybridConfig(tasks=case["tasks"], checkpoint_dir=str(ck), base_model_path=str(tmp_path / "base.pt"),
                          svd_store_artifacts=True, svd_max_rank=64, output_dir=str(tmp_path / "out"),
                          artifact_dir=str(tmp_path / "art"), device="cuda", **case["config"])
    res = run_svd_hybrid_pipeline(cfg, verbose=False)
    out = reconstruct_from_artifacts(str(tmp_path / "art"), str(tmp_path / "base.pt"), str(tmp_path / "re.pt"), "cpu")
    for p, m in res["merged_state_dict"].items():
        assert torch.allclose(out["merged_state_dict"][p], m.cpu(), rtol=1e-4, atol=1e-7), p
    assert os.path.exists(tmp_path / "re.pt")


# ---- world-size invariance of the parameter-sharded path -----------------------------------------------------
@pytest.mark.parametrize("weighting,noise,n_tasks", [("uniform", False, 8), ("cluster", False, 8), ("cluster", True, 8),
                                                     ("uniform", False, 20)])
def test_sharded_results_bit_identical_for_any_world_size(cuda_device, weighting, noise, n_tasks):
    """Parameters are independent units: processing the shards of world size 2/4/8 (logical ranks, one
    after the other on this GPU) must reproduce the single-rank result bit for bit."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    tasks = synth.task_names(n_tasks)
    shapes = dict(parity.MEDIUM_SHAPES)
    shapes.update({f"extra{i}.weight": (64 + i, 33) for i in range(9)})
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=31)
    masks = synth.make_masks(shapes, tasks, 0.6, seed=32)
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_weighting=weighting, svd_store_artifacts=False,
                          svd_include_noise=noise)
    full = MergeJob(base, fts, masks, cfg, "cuda").run()
    merged_full = full.merged_state_dict()
    diag_full = full.results()["diagnostics"]["per_parameter"]
    gram = torch.from_numpy(full.whole_model_gram).cuda().view(-1) if weighting == "cluster" else None
    cost = {k: int(np.prod(v)) * (n_tasks + 1) for k, v in shapes.items()}
    for world in (2, 4, 8):
        owner = sharding.lpt_partition(cost, world)
        seen = set()
        for rank in range(world):
            mine = [n for n, r in owner.items() if r == rank]
            job = MergeJob(base, fts, masks, cfg, "cuda", param_filter=mine)
            if gram is not None:
                job.gram_reduce_hook = lambda g_local: gram.clone()      # what the NCCL all-reduce would deliver
            job.run()
            part = job.merged_state_dict()
            assert sorted(part.keys()) == sorted(mine)
            d = job.results()["diagnostics"]["per_parameter"]
            for n in mine:
                # bit patterns, so that NaNs (degenerate few-row noise regions quantise to NaN like in the reference)
                # count as equal
                assert torch.equal(part[n].view(torch.int32), merged_full[n].view(torch.int32)), (world, rank, n)
                if n in diag_full:
                    assert d[n] == diag_full[n], (world, rank, n)
            seen.update(mine)
        assert seen == set(shapes)


# ---- host masks travel bit-packed; device masks stay torch.bool bytes: same results, bit for bit ---------------
@pytest.mark.parametrize("strategy,n_tasks,dtype,noise", [("union", 3, torch.float32, False),
                                                          ("majority", 8, torch.float32, True),
                                                          ("intersection", 8, torch.bfloat16, False),
                                                          ("majority", 12, torch.float32, False),
                                                          ("union", 16, torch.float16, False)])
def test_host_bit_packed_masks_equal_device_byte_masks(cuda_device, strategy, n_tasks, dtype, noise):
    from svd_quantization_task_merging_b200.engine import MergeJob, pack_state_dict
    tasks = synth.task_names(n_tasks)
    shapes = dict(parity.MEDIUM_SHAPES)
    shapes.update({"odd.weight": (1031, 7), "tiny.bias": (5,), "chunk.weight": (1024, 3), "word.weight": (33,)})
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=51)
    base = {k: v.to(dtype) for k, v in base.items()}
    fts = {t: {k: v.to(dtype) for k, v in sd.items()} for t, sd in fts.items()}
    masks = synth.make_masks(shapes, tasks, 0.55, seed=52)
    del masks[tasks[0]]["odd.weight"]                         # a task without a mask for one parameter
    masks[tasks[1]]["tiny.bias"] = torch.ones(6, dtype=torch.bool)        # wrong shape: parameter runs unmasked
    masks[tasks[-1]] = {k: v.to(torch.uint8) * 3 for k, v in masks[tasks[-1]].items()}   # non-bool mask dtype
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy=strategy, svd_store_artifacts=False,
                          svd_include_noise=noise, svd_weighting="cluster")
    variants = {
        "device_bytes": {t: {k: v.cuda() for k, v in m.items()} for t, m in masks.items()},
        "host_dicts": masks,
        "host_packed": {t: (pack_state_dict(m, pin=True) if m and len({v.dtype for v in m.values()}) == 1 else m)
                        for t, m in masks.items()},
    }
    results = {}
    for tag, mk in variants.items():
        job = MergeJob(base, fts, mk, cfg, "cuda")
        assert job.mask_bits == (tag != "device_bytes")
        job.run()
        results[tag] = (job.merged_state_dict(), job.combined_masks(), job.results()["diagnostics"]["per_parameter"])
    ref_m, ref_c, ref_d = results["device_bytes"]
    itype = torch.int32 if dtype == torch.float32 else torch.int16
    for tag in ("host_dicts", "host_packed"):
        m, c, d = results[tag]
        assert sorted(c.keys()) == sorted(ref_c.keys())
        for k in ref_c:
            assert torch.equal(c[k], ref_c[k]), (tag, k)
        for k in ref_m:
            assert torch.equal(m[k].view(itype), ref_m[k].view(itype)), (tag, k)
        # through json so that NaNs (few-task degenerate blocks quantise to NaN like in the reference) compare equal
        canon = lambda x: json.dumps(x, sort_keys=True, default=lambda o: o.item() if hasattr(o, "item") else str(o))
        assert canon(d) == canon(ref_d), tag


# ---- two-level Gram reduction of parameters with >= 512 tiles --------------------------------------------------
@pytest.mark.parametrize("weighting,noise", [("uniform", False), ("cluster", False), ("uniform", True)])
def test_two_level_gram_reduction_of_many_tile_parameters(cuda_device, weighting, noise):
    """tile_elems = 1024 makes 512+ tiles cheap: 511 tiles (one level), 512 (64 ranges of 8), 513 (57 ranges of 9),
    684 and 2000 tiles.  The reduced Grams / counts match an fp64 torch Gram of the masked task vectors, and a job
    that owns only some of the parameters reproduces them bit for bit."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    tasks = synth.task_names(8)
    shapes = {"t511.weight": (511 * 1024 - 3,), "t512.weight": (512, 1024), "t513.weight": (512 * 1024 + 1,),
              "t684.weight": (700_000,), "t2000.weight": (2000, 1024), "small.weight": (300, 7)}
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=61, device="cuda")
    masks = synth.make_masks(shapes, tasks, 0.5, seed=62, device="cuda")
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy="majority", svd_weighting=weighting,
                          svd_store_artifacts=False, svd_include_noise=noise)
    job = MergeJob(base, fts, masks, cfg, "cuda", tile_elems=1024).run()
    g = job.groups[torch.float32]
    got = g.t["gram_masked"].view(-1, 8, 8).clone()
    got_n = g.tn["gram"].view(-1, 8, 8).clone() if noise else None
    dm = g.t["dm"].clone()
    for i, k in enumerate(g.names):
        maj = (2 * torch.stack([masks[t][k] for t in tasks]).sum(0) >= 8).view(-1)
        T = torch.stack([(fts[t][k] - base[k]).view(-1) for t in tasks], 1).double()
        ref = T[maj].T @ T[maj]
        assert int(dm[i]) == int(maj.sum())
        assert (got[i] - ref).abs().max() <= 2e-6 * ref.abs().max(), k
        if noise:
            ref_n = T[~maj].T @ T[~maj]
            assert (got_n[i] - ref_n).abs().max() <= 2e-6 * ref_n.abs().max(), k
    if weighting == "cluster":
        T_all = torch.cat([torch.stack([(fts[t][k] - base[k]).view(-1) for t in tasks], 1).double() for k in g.names])
        whole = torch.from_numpy(job.whole_model_gram).cuda()
        ref = T_all.T @ T_all
        assert (whole - ref).abs().max() <= 2e-6 * ref.abs().max()
    merged = job.merged_state_dict()
    part = MergeJob(base, fts, masks, cfg, "cuda", tile_elems=1024, param_filter=["t513.weight", "t2000.weight"])
    if weighting == "cluster":
        gram = torch.from_numpy(job.whole_model_gram).cuda().view(-1)
        part.gram_reduce_hook = lambda g_local: gram.clone()
    part.run()
    gp = part.groups[torch.float32]
    for i, k in enumerate(gp.names):
        j = g.names.index(k)
        assert torch.equal(gp.t["gram_masked"].view(-1, 8, 8)[i], got[j]), k
        assert torch.equal(part.merged_state_dict()[k].view(torch.int32), merged[k].view(torch.int32)), k


def test_reload_with_noise_region_and_batched_merge_matches_per_parameter(cuda_device, tmp_path):
    """Artifacts of a masked run with svd_include_noise: the batched reload merge (K11, one launch over the stored
    bases) reproduces the pipeline's merged model and equals the per-parameter operator merge_parameter."""
    from src.svd_hybrid.cli import run_svd_hybrid_pipeline
    from src.svd_hybrid.reload import reconstruct_from_artifacts
    from src.svd_hybrid.merge import merge_all_parameters, merge_parameter
    from src.svd_hybrid.rtvq import RTVQQuantizer
    from src.svd_hybrid.storage import load_all_artifacts, load_combined_masks
    case = _gold("majority_noise_uniform")
    ck, md = _write_case(tmp_path, case)
    cfg = SVDHybridConfig(tasks=case["tasks"], checkpoint_dir=str(ck), base_model_path=str(tmp_path / "base.pt"),
                          mask_dir=str(md), svd_store_artifacts=True, svd_max_rank=64, output_dir=str(tmp_path / "out"),
                          artifact_dir=str(tmp_path / "art"), device="cuda", **case["config"])
    res = run_svd_hybrid_pipeline(cfg, verbose=False)
    out = reconstruct_from_artifacts(str(tmp_path / "art"), str(tmp_path / "base.pt"), str(tmp_path / "re.pt"), "cpu")
    for p, m in res["merged_state_dict"].items():
        assert torch.allclose(out["merged_state_dict"][p], m.cpu(), rtol=1e-4, atol=1e-6), p
    art = load_all_artifacts(str(tmp_path / "art"), device="cpu")
    masks = load_combined_masks(str(tmp_path / "art"), device="cpu")
    weights = art["diagnostics"]["task_weights"]
    shapes = {p: torch.Size(d["original_shape"]) for p, d in art["diagnostics"]["per_parameter"].items()}
    batched = merge_all_parameters(art["compressed"], art["bases"], masks, weights, shapes, art["config"], "cpu", verbose=False)
    q = RTVQQuantizer(num_bits=art["config"].svd_low_bits, num_stages=art["config"].svd_rtvq_stages)
    for p in batched:
        one = merge_parameter(p, art["compressed"][p], art["bases"][p], weights, q, shapes[p], mask=masks.get(p),
                              include_noise=True, noise_shrink=art["config"].svd_noise_shrink, device="cpu")
        assert torch.allclose(batched[p], one, rtol=1e-5, atol=1e-8), p


def test_k14_projection_and_expansion(cuda_device):
    """project_to_basis / reconstruct_from_coefficients on K14 (no library matmul): against fp64 torch on fp16 and
    fp32 bases, strided column slices, more than 32 columns, with and without a mean; launch-independent results."""
    from svd_quantization_task_merging_b200.svd_hybrid import _ops
    from src.svd_hybrid.compress import project_to_basis
    from src.svd_hybrid.merge import reconstruct_from_coefficients
    g = torch.Generator().manual_seed(5)
    for rows, cols, dt in ((1, 1, torch.float32), (257, 3, torch.float16), (70001, 8, torch.float16),
                           (300000, 20, torch.float32), (5000, 50, torch.float32), (1000003, 32, torch.float16)):
        U = (torch.randn(rows, cols, generator=g) / rows ** 0.5).to(dt)
        x = torch.randn(rows, generator=g)
        m = torch.randn(rows, 1, generator=g) * 0.1
        k = max(1, cols // 3) if cols > 1 else 1
        Uh, Ul = U[:, :k], U[:, k:]                              # column slices: leading dimension = cols
        ch, cl = project_to_basis(x, Uh, Ul)
        ref = U.double().T @ x.double()
        scale = (U.double().abs().T @ x.double().abs()).clamp_min(1e-30)
        assert ch.dtype == torch.float32 and ch.device.type == "cpu" and ch.shape == (k,) and cl.shape == (cols - k,)
        assert ((torch.cat([ch, cl]).double() - ref).abs() / scale).max().item() < 2e-6
        chm, clm = project_to_basis(x, Uh, Ul, mean=m)
        refm = U.double().T @ (x.double() - m.squeeze(1).double())
        scale_m = (U.double().abs().T @ (x.double() - m.squeeze(1).double()).abs()).clamp_min(1e-30)
        assert ((torch.cat([chm, clm]).double() - refm).abs() / scale_m).max().item() < 2e-6
        again = _ops.project(x.cuda(), U.cuda())
        assert torch.equal(again.cpu(), torch.cat([ch, cl]))                  # fixed reduction order
        c = torch.randn(cols, generator=g)
        rec = reconstruct_from_coefficients(c[:k], c[k:], Uh, Ul, "cpu", mean=m)
        want = U.double() @ c.double() + m.squeeze(1).double()
        bound = U.double().abs() @ c.double().abs() + m.squeeze(1).double().abs()
        assert rec.dtype == torch.float32 and rec.shape == (rows,)
        assert ((rec.double() - want).abs() / bound.clamp_min(1e-30)).max().item() < 2e-6
        rec0 = reconstruct_from_coefficients(c[:k], c[k:], Uh, Ul, "cuda")
        assert rec0.is_cuda
        assert ((rec0.cpu().double() - U.double() @ c.double()).abs() / bound.clamp_min(1e-30)).max().item() < 2e-6
    with pytest.raises(ValueError):
        _ops.project(torch.zeros(5), torch.zeros(6, 2))


def test_k14_mask_select_and_scatter(cuda_device):
    """apply_mask_to_tensor / get_unmasked_portion / reconstruct_from_masked on the K14 selection kernels: bit-equal
    to torch boolean indexing for every element size, ragged sizes, unaligned mask storage, empty / full masks."""
    from src.svd_hybrid.mask_loader import apply_mask_to_tensor, get_unmasked_portion, reconstruct_from_masked
    g = torch.Generator().manual_seed(9)
    for n in (1, 15, 16, 4095, 4096, 4097, 100003, 3 * 4096 * 257 + 5):
        for dt in (torch.float32, torch.float16, torch.float64, torch.uint8, torch.bfloat16):
            if n > 200000 and dt not in (torch.float32, torch.float16):
                continue
            x = (torch.randn(n, generator=g) * 50).to(dt)
            for p in (0.0, 0.3, 1.0):
                store = torch.rand(n + 3, generator=g) < p
                mask = store[3:] if n % 2 else store[:n]          # odd sizes: mask storage off 16-byte alignment
                kept = apply_mask_to_tensor(x, mask)
                rest = get_unmasked_portion(x, mask)
                assert kept.dtype == dt and kept.device.type == "cpu"
                assert torch.equal(kept, x[mask]) and torch.equal(rest, x[~mask])
                back = reconstruct_from_masked(kept, rest, mask, x.shape)
                assert torch.equal(back, x)
                only = reconstruct_from_masked(kept, None, mask, x.shape)
                assert torch.equal(only, torch.where(mask, x, torch.zeros_like(x)))
    x = torch.randn(6, 7, 11, generator=g).cuda()
    mask = (torch.rand(6, 7, 11, generator=g) < 0.5).cuda()
    assert apply_mask_to_tensor(x, mask).is_cuda and torch.equal(apply_mask_to_tensor(x, mask), x.flatten()[mask.flatten()])
    assert torch.equal(reconstruct_from_masked(apply_mask_to_tensor(x, mask), get_unmasked_portion(x, mask), mask, x.shape), x)
    with pytest.raises(ValueError):
        apply_mask_to_tensor(torch.zeros(4), torch.zeros(5, dtype=torch.bool))
    with pytest.raises(ValueError):
        reconstruct_from_masked(torch.zeros(3), None, torch.ones(5, dtype=torch.bool), torch.Size([5]))


def test_pipeline_device_cpu_means_host_resident_results(cuda_device, tmp_path, capsys):
    """The reference's integration tests configure device="cpu" (tests/test_integration.py:77,165).  Here the arithmetic
    runs on the GPU regardless; the setting decides where the returned state dict lives, and the run says so."""
    from src.svd_hybrid.cli import run_svd_hybrid_pipeline
    case = _gold("union_uniform")
    ck, md = _write_case(tmp_path, case)
    out = {}
    for dev in ("cpu", "cuda"):
        cfg = SVDHybridConfig(tasks=case["tasks"], checkpoint_dir=str(ck), base_model_path=str(tmp_path / "base.pt"),
                              mask_dir=str(md), svd_store_artifacts=False, svd_eval_reconstruction=True, svd_max_rank=64,
                              output_dir=str(tmp_path / f"out_{dev}"), artifact_dir=str(tmp_path / f"art_{dev}"),
                              device=dev, **case["config"])
        out[dev] = run_svd_hybrid_pipeline(cfg, verbose=False)
        text = capsys.readouterr().out
        assert ('device="cpu" requested' in text) == (dev == "cpu")
    for name, t in out["cpu"]["merged_state_dict"].items():
        assert t.device.type == "cpu"
        assert torch.equal(t, out["cuda"]["merged_state_dict"][name].cpu()), name
    assert out["cpu"]["diagnostics"]["task_weights"] == out["cuda"]["diagnostics"]["task_weights"]


def test_reload_script_rebuilds_and_verifies(cuda_device, tmp_path):
    """scripts/reload_svd_hybrid.py (reference scripts/reload_svd_hybrid.py:149-256): rebuild from the artifact
    directory, save, and --verify against the merged model of the original run."""
    import shutil
    import subprocess
    import sys
    from src.svd_hybrid.cli import run_svd_hybrid_pipeline
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "scripts", "reload_svd_hybrid.py")
    case = _gold("union_uniform")
    ck, md = _write_case(tmp_path, case)
    cfg = SVDHybridConfig(tasks=case["tasks"], checkpoint_dir=str(ck), base_model_path=str(tmp_path / "base.pt"),
                          mask_dir=str(md), svd_store_artifacts=True, svd_max_rank=64, output_dir=str(tmp_path / "out"),
                          artifact_dir=str(tmp_path / "art"), device="cuda", **case["config"])
    res = run_svd_hybrid_pipeline(cfg, verbose=False)

    def run(*args):
        return subprocess.run([sys.executable, script, *args], cwd=root, capture_output=True, text=True, timeout=600)
    # 1. no merged model inside the artifact directory: reconstructed from bases + codes through the batched reload merge
    r = run("--artifact-dir", str(tmp_path / "art"), "--output-path", str(tmp_path / "re" / "model.pt"), "--verbose")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Successfully reloaded model" in r.stdout and "Reload Complete!" in r.stdout
    re_sd = torch.load(tmp_path / "re" / "model.pt", weights_only=False)
    for p, m in res["merged_state_dict"].items():
        assert torch.allclose(re_sd[p], m.cpu(), rtol=1e-4, atol=1e-6), p
    # 2. --verify needs --merged-model-path
    r = run("--artifact-dir", str(tmp_path / "art"), "--verify")
    assert r.returncode == 1 and "--merged-model-path required" in r.stdout
    # 3. with the run's merged model stored next to the artifacts the reload is that file and the checksums agree
    shutil.copy(tmp_path / "out" / "merged_state_dict.pt", tmp_path / "art" / "merged_state_dict.pt")
    r = run("--artifact-dir", str(tmp_path / "art"), "--verify", "--merged-model-path",
            str(tmp_path / "out" / "merged_state_dict.pt"), "--verbose")
    assert r.returncode == 0 and "Verification PASSED" in r.stdout and "Checksums match: True" in r.stdout, r.stdout
    # 4. a missing artifact directory is reported, exit code 1
    r = run("--artifact-dir", str(tmp_path / "nope"))
    assert r.returncode == 1 and "Artifact directory not found" in r.stdout


@pytest.mark.parametrize("name", ["union_uniform", "majority_performance_3stage", "nomask_fp32_nocenter",
                                  "majority_noise_uniform", "wide20_union_uniform"])
def test_reload_of_reference_written_artifacts(cuda_device, tmp_path, name):
    """Format compatibility of the reload path: the bases, coefficient objects and diagnostics the REAL reference
    produced (golden fixture) are written in its artifact layout and re-merged by the batched reload kernel (K11).
    With the reference's own U there is no sign freedom left, so the result must equal the reference's merged model
    to fp32 round-off (reference reload.py:142-238, merge.py:304-426; the combined masks come from the additive
    combined_masks.pt, which the reference's reload lacks)."""
    from src.svd_hybrid.reload import reconstruct_from_artifacts
    from src.svd_hybrid.storage import save_all_artifacts, save_combined_masks
    case = _gold(name)
    torch.save(dict(case["base"]), tmp_path / "base.pt")
    cfg = SVDHybridConfig(tasks=case["tasks"], base_model_path=str(tmp_path / "base.pt"), svd_max_rank=64,
                          svd_store_artifacts=True, **case["config"])
    art = str(tmp_path / "art")
    save_all_artifacts(case["bases"], case["compressed"], case["diagnostics"], cfg, art)
    if case["combined_masks"]:
        save_combined_masks(case["combined_masks"], art)
    out = reconstruct_from_artifacts(art, str(tmp_path / "base.pt"), str(tmp_path / "re.pt"), "cpu")
    worst = 0.0
    for p, gm in case["merged_state_dict"].items():
        mine = out["merged_state_dict"][p]
        assert mine.shape == gm.shape and mine.dtype == gm.dtype
        den = (gm - case["base"][p]).norm().item()
        err = (mine - gm).norm().item() / max(den, 1e-30)
        worst = max(worst, err)
        assert err < 2e-6, (p, err)
    print(name, "reload of reference artifacts: worst rel L2 of the merged delta", worst)


@pytest.mark.parametrize("name", ["union_uniform", "majority_performance_3stage", "nomask_fp32_nocenter",
                                  "wide20_union_uniform"])
def test_operator_api_on_reference_bases(cuda_device, name):
    """The operator-by-operator mirrors (K14 projection / expansion / mask selection, K4 RTVQ) on the REAL reference's
    stored bases: compress_parameter must reproduce the reference's own fp16 coefficients and RTVQ codes, and
    compute_parameter_diagnostics its per-task reconstruction errors -- no sign freedom, the basis is the reference's
    (compress.py:114-170, diagnostics.py:120-231)."""
    from src.svd_hybrid.compress import compress_parameter
    from src.svd_hybrid.diagnostics import compute_parameter_diagnostics
    from src.svd_hybrid.rtvq import RTVQQuantizer
    case = _gold(name)
    cfg = SVDHybridConfig(tasks=case["tasks"], svd_max_rank=64, **case["config"])
    quant = RTVQQuantizer(num_bits=cfg.svd_low_bits, num_stages=cfg.svd_rtvq_stages)
    tvs = {t: {p: case["finetuned"][t][p] - case["base"][p] for p in case["base"]} for t in case["tasks"]}
    n_hi = eq_hi = n_code = eq_code = 0
    for p, basis in case["bases"].items():
        mask = case["combined_masks"].get(p) if case["combined_masks"] else None
        mine = compress_parameter(p, tvs, mask, basis, quant, include_noise=False, min_mask_size=cfg.svd_min_mask_size)
        for t in case["tasks"]:
            g, m = case["compressed"][p][t]["masked"], mine[t]["masked"]
            assert m["c_high_fp16"].dtype == torch.float16 and m["c_high_fp16"].shape == g["c_high_fp16"].shape
            # fp16 values at most one ulp apart (fp32 summation order of U^T d differs from the reference's BLAS)
            a, b = m["c_high_fp16"].view(torch.int16).int(), g["c_high_fp16"].view(torch.int16).int()
            assert (a - b).abs().max().item() <= 1, (p, t)
            n_hi += a.numel()
            eq_hi += int((a == b).sum())
            for x, y in zip(m["c_low_quant"]["payloads"], g["c_low_quant"]["payloads"]):
                n_code += x["quantized"].numel()
                eq_code += int((x["quantized"] == y["quantized"]).sum())
        gd = case["diagnostics"]["per_parameter"][p]
        d = compute_parameter_diagnostics(p, tvs, case["compressed"][p], basis, mask, quant)
        assert d["masked_size"] == gd["masked_size"] and d["basis"]["k"] == gd["basis"]["k"]
        for t in case["tasks"]:
            for key in ("absolute_error", "relative_error", "original_norm", "reconstructed_norm", "mean_absolute_error"):
                assert d["reconstruction_errors"][t][key] == pytest.approx(gd["reconstruction_errors"][t][key], rel=2e-5), (p, t, key)
            assert d["compression_ratios"][t] == gd["compression_ratios"][t]
    print(name, f"c_high fp16 bit-equal {eq_hi}/{n_hi}, RTVQ codes equal {eq_code}/{n_code}")
    assert eq_hi >= 0.99 * n_hi and eq_code >= 0.99 * n_code         # observed on B200: all of them, in all four cases
