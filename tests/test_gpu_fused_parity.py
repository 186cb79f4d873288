"""GPU parity of the fused SVD-Hybrid merge (K1 -> K2 -> K3 through the C ABI) against the oracle
on the same seeded inputs.  Tolerances are the ones written in tests/parity.py."""
import numpy as np
import pytest
import torch

from svd_quantization_task_merging_b200 import synth
from tests import parity

pytestmark = pytest.mark.gpu


def _summary(rep):
    return (f"params={rep['params']} codes {rep['code_equal']}/{rep['code_total']} "
            f"c_high {rep['chigh_equal']}/{rep['chigh_total']} max_merged_rel={rep['max_merged_rel']:.2e} "
            f"flipped={rep['flipped_params']}")


@pytest.mark.parametrize("strategy,mask_p", [("union", 0.3), ("intersection", 0.9), ("majority", 0.5), ("union", None)])
def test_medium_shapes_masks(cuda_device, strategy, mask_p):
    """configs[0]-style: 8 tasks, energy 0.9, 4-bit x 2 stages, uniform weights; all three tall-mask rules."""
    ref, res, _ = parity.run_both(parity.MEDIUM_SHAPES, 8, mask_p=mask_p, svd_mask_strategy=strategy,
                                  svd_energy_threshold=0.9, svd_low_bits=4, svd_rtvq_stages=2)
    rep = parity.compare_run(ref, res)
    print(_summary(rep))
    assert rep["code_equal"] >= 0.99 * rep["code_total"]          # observed 0.995 .. 1.0 (profiles/parity_rates.json)


def test_performance_weighting_three_stages(cuda_device):
    """configs[1]-style: majority mask, performance-softmax weights (T=5), 4-bit x 3 stages."""
    ref, res, _ = parity.run_both(parity.MEDIUM_SHAPES, 8, mask_p=0.5, svd_mask_strategy="majority",
                                  svd_weighting="performance", svd_weighting_temperature=5.0,
                                  svd_energy_threshold=0.9, svd_rtvq_stages=3)
    rep = parity.compare_run(ref, res)
    print(_summary(rep))
    w_ref, w_new = ref["weights"], res["diagnostics"]["task_weights"]
    assert w_ref == w_new


def test_cluster_weighting(cuda_device):
    """configs[2]-style: intersection mask, cluster weighting (k=2)."""
    ref, res, _ = parity.run_both(parity.MEDIUM_SHAPES, 8, mask_p=0.9, svd_mask_strategy="intersection",
                                  svd_weighting="cluster", svd_cluster_k=2, svd_energy_threshold=0.9)
    rep = parity.compare_run(ref, res)
    print(_summary(rep))
    assert res["diagnostics"]["cluster_assignments"] == ref["cluster_assignments"]


def test_cluster_partition_from_gram_matches_full_kmeans(cuda_device):
    """The k-means partition computed from the K1 whole-model Gram equals the reference's
    full-feature k-means (clustering.py:198-245) on the same task vectors."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig
    from oracle import svd_hybrid_ref as R
    tasks = synth.task_names(8)
    base, fts = synth.make_checkpoints(parity.MEDIUM_SHAPES, tasks, family="parity", seed=77)
    masks = synth.make_masks(parity.MEDIUM_SHAPES, tasks, 0.9, seed=78)
    cfg = SVDHybridConfig(tasks=tasks, svd_weighting="cluster", svd_mask_strategy="intersection",
                          svd_energy_threshold=0.9, svd_store_artifacts=False)
    job = MergeJob(base, fts, masks, cfg, "cuda").run()      # default cluster backend
    tvs = {t: R.task_vector(base, fts[t]) for t in tasks}
    full = R.cluster_tasks_full(tvs, 2)
    mine = job.cluster_assignments
    assert all((full[a] == full[b]) == (mine[a] == mine[b]) for a in tasks for b in tasks)
    # and the Gram itself against fp64 numpy
    X = np.stack([torch.cat([tvs[t][p].flatten() for p in sorted(tvs[t])]).double().numpy() for t in tasks])
    G = X @ X.T
    assert np.abs(job.whole_model_gram - G).max() <= 1e-6 * np.abs(G).max()


def test_fp32_basis_no_center(cuda_device):
    ref, res, _ = parity.run_both(parity.MEDIUM_SHAPES, 6, mask_p=0.5, svd_mask_strategy="union", svd_fp16=False,
                                  svd_center=False, svd_energy_threshold=0.8, svd_rtvq_stages=2)
    rep = parity.compare_run(ref, res)
    print(_summary(rep))


@pytest.mark.parametrize("bits,stages", [(2, 1), (2, 4), (4, 1), (8, 2), (3, 3)])
def test_rtvq_sweep_through_pipeline(cuda_device, bits, stages):
    """configs[3]-style RTVQ sweep (bits x stages) through the fused path."""
    ref, res, _ = parity.run_both(synth.toy_shapes(), 8, svd_energy_threshold=0.9, svd_low_bits=bits,
                                  svd_rtvq_stages=stages)
    rep = parity.compare_run(ref, res)
    print(_summary(rep))


def test_toy_model_of_reference_integration_test(cuda_device):
    """tests/test_integration.py:12-19 shapes, 4 tasks, no masks."""
    ref, res, _ = parity.run_both(synth.toy_shapes(), 4, svd_energy_threshold=0.8, svd_fp16=False)
    parity.compare_run(ref, res)


def test_degenerate_iid_inputs_reproduce_reference_nans(cuda_device):
    """iid Gaussian task vectors, centred: k = N-1, a 1-element low block, scale = inf -> every merged
    masked element is NaN in the reference (rtvq.py:17).  Same NaN positions here."""
    ref, res, _ = parity.run_both(synth.toy_shapes(), 8, family="throughput", svd_energy_threshold=0.95,
                                  sign_align=True)
    for name, m in ref["merged_state_dict"].items():
        assert parity.nan_positions_equal(res["merged_state_dict"][name], m), name
    assert any(torch.isnan(m).any() for m in ref["merged_state_dict"].values())


def test_small_mask_is_skipped(cuda_device):
    """mask.sum() < svd_min_mask_size -> no basis, parameter stays equal to base (cli.py:332)."""
    tasks = synth.task_names(4)
    shapes = {"w": (64, 64), "v": (512,)}
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=3)
    masks = {t: {"w": torch.zeros(64, 64, dtype=torch.bool), "v": torch.ones(512, dtype=torch.bool)} for t in tasks}
    for t in tasks:
        masks[t]["w"][0, :5] = True            # 5 < 10 masked elements
    from oracle import svd_hybrid_ref as R
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    ref_cfg, cfg = parity.make_cfgs(tasks, svd_energy_threshold=0.8)
    ref = R.run_reference_path(base, fts, masks, ref_cfg)
    ref["_base"] = base
    res = merge_state_dicts(base, fts, masks, cfg, "cuda", sign_ref={p: b["Vh"] for p, b in ref["bases"].items()})
    assert "w" not in res["bases"] and "w" not in ref["bases"]
    assert torch.equal(res["merged_state_dict"]["w"].cpu(), base["w"])
    parity.compare_run(ref, res)


def test_missing_task_parameter_and_partial_masks(cuda_device):
    """A task that lacks a parameter, and a task without a mask for a parameter."""
    tasks = synth.task_names(5)
    shapes = {"a": (90, 41), "b": (3000,)}
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=9)
    masks = synth.make_masks(shapes, tasks, 0.6, seed=10)
    del fts[tasks[2]]["b"]
    del masks[tasks[1]]["a"]
    masks[tasks[4]] = None
    from oracle import svd_hybrid_ref as R
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    ref_cfg, cfg = parity.make_cfgs(tasks, svd_energy_threshold=0.85, svd_mask_strategy="majority")
    ref = R.run_reference_path(base, fts, masks, ref_cfg)
    ref["_base"] = base
    # the oracle's Vh has one column per ACTIVE task: scatter it to task positions
    sign_ref = {}
    for p, b in ref["bases"].items():
        active = [i for i, t in enumerate(tasks) if p in fts[t]]
        vh = torch.zeros(len(tasks), len(tasks), dtype=torch.float64)
        vh[: b["Vh"].shape[0], active] = b["Vh"].double()
        sign_ref[p] = vh
    res = merge_state_dicts(base, fts, masks, cfg, "cuda", sign_ref=sign_ref)
    rep = parity.compare_run(ref, res)
    print(_summary(rep))


@pytest.mark.parametrize("n_tasks,noise", [(6, False), (6, True), (18, False)])
def test_cluster_weighting_with_missing_parameters(cuda_device, n_tasks, noise):
    """merge_with_clustering (merge.py:586-626) renormalises the member weights inside each cluster over the members
    that HAVE the parameter, and a cluster none of whose members has it contributes zeros -- mean included.
    Parameter "b" is missing in one member of cluster 0; parameter "c" in ALL members of cluster 1."""
    from oracle import svd_hybrid_ref as R
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    tasks = synth.task_names(n_tasks)
    shapes = {"a": (90, 41), "b": (3000,), "c": (64, 70)}
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=19)
    masks = synth.make_masks(shapes, tasks, 0.6, seed=20)
    assign = {t: (0 if i % 3 else 1) for i, t in enumerate(tasks)}          # cluster 1: tasks 0, 3, ...
    cl0 = [t for t in tasks if assign[t] == 0]
    del fts[cl0[1]]["b"]
    for t in tasks:
        if assign[t] == 1:
            del fts[t]["c"]
    ref_cfg, cfg = parity.make_cfgs(tasks, svd_energy_threshold=0.85, svd_mask_strategy="union", svd_weighting="cluster",
                                    svd_include_noise=noise)
    ref = R.run_reference_path(base, fts, masks, ref_cfg, assignments=assign)
    ref["_base"] = base

    def scatter(bases):
        out = {}
        for p, b in bases.items():
            active = [i for i, t in enumerate(tasks) if p in fts[t]]
            vh = torch.zeros(len(tasks), len(tasks), dtype=torch.float64)
            vh[: b["Vh"].shape[0], active] = b["Vh"].double()
            out[p] = vh
        return out
    res = merge_state_dicts(base, fts, masks, cfg, "cuda", sign_ref=scatter(ref["bases"]),
                            sign_ref_noise=scatter(ref.get("bases_noise") or {}), cluster_assignments=assign)
    rep = parity.compare_run(ref, res)
    print(_summary(rep))
    # the quirk is really exercised: the reference's "c" differs from an all-clusters-present merge of the same inputs
    scal = res["job"]._fetch()[torch.float32]["scal"]
    names = res["job"].groups[torch.float32].names
    assert scal[names.index("c"), 2] < 0.999 and abs(scal[names.index("a"), 2] - 1.0) < 1e-6


def test_sharded_upload_moves_only_the_shard(cuda_device):
    """Host (packed, pinned) state dicts + param_filter: only the owned tensors cross PCIe, and the result equals
    the unfiltered merge bit for bit; a rank with an EMPTY shard still runs (and would join the collectives)."""
    from svd_quantization_task_merging_b200 import sharding
    from svd_quantization_task_merging_b200.engine import MergeJob, merge_state_dicts, pack_state_dict
    tasks = synth.task_names(8)
    shapes = dict(parity.MEDIUM_SHAPES)
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=41)
    masks = synth.make_masks(shapes, tasks, 0.7, seed=42)
    _, cfg = parity.make_cfgs(tasks, svd_energy_threshold=0.9, svd_mask_strategy="intersection", svd_weighting="cluster")
    assign = {t: i % 2 for i, t in enumerate(tasks)}
    full = merge_state_dicts(base, fts, masks, cfg, "cuda", cluster_assignments=assign)
    hb = pack_state_dict(base, pin=True)
    hf = {t: pack_state_dict(fts[t], pin=True) for t in tasks}
    hm = {t: pack_state_dict(masks[t], pin=True) for t in tasks}
    whole = MergeJob(hb, hf, hm, cfg, "cuda", cluster_assignments=assign).h2d_bytes
    cost = {k: int(np.prod(v)) * 9 for k, v in shapes.items()}
    world = 3
    owner = sharding.lpt_partition(cost, world)
    moved = 0
    for r in range(world):
        mine = [n for n, o in owner.items() if o == r]
        job = MergeJob(hb, hf, hm, cfg, "cuda", cluster_assignments=assign, param_filter=mine)
        moved += job.h2d_bytes
        job.run()
        got = job.merged_state_dict()
        assert sorted(got) == sorted(mine)
        for n in mine:
            assert torch.equal(got[n], full["merged_state_dict"][n]), n
        # upload accounting: this shard's tensors (+ alignment padding) and its bit-packed masks, nothing else
        own = sum(int(np.prod(shapes[n])) for n in mine)
        bound = own * 4 * 9 + own * 8 // 8 + 4096 * len(mine) * 17
        assert job.h2d_bytes <= bound, (r, job.h2d_bytes, bound)
    assert moved <= whole * 1.02 + 65536, (moved, whole)
    empty = MergeJob(hb, hf, hm, cfg, "cuda", param_filter=[])
    empty.gram_reduce_hook = lambda g: g
    empty.run()
    assert len(empty.merged_state_dict()) == 0 and empty.h2d_bytes <= 4096, empty.h2d_bytes


def test_without_sign_hint_matches_at_quantisation_noise_level(cuda_device):
    """Production mode (no sign hint): singular-vector signs are a free choice, RTVQ is not sign-symmetric,
    so the merged weights agree with the oracle only at the quantisation-noise level."""
    ref, res, _ = parity.run_both(parity.MEDIUM_SHAPES, 8, mask_p=0.3, svd_energy_threshold=0.9, sign_align=False)
    for name, d_ref in ref["merged_deltas"].items():
        m_new = res["merged_state_dict"][name].cpu()
        d_new = m_new - ref["_base"][name]
        if torch.isfinite(d_ref).all() and d_ref.norm() > 0:
            assert parity.rel_l2(d_new, d_ref) < 5e-2, name


def test_bf16_inputs(cuda_device):
    """bf16 checkpoints (configs[4] dtype): delta is formed in bf16 like `ft - base` on bf16 tensors;
    the reference's own SVD rejects bf16 on CPU, so the oracle runs on the bf16-rounded deltas in fp32."""
    from oracle import svd_hybrid_ref as R
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    tasks = synth.task_names(8)
    shapes = {"w": (256, 130), "b": (1000,)}
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=21, dtype=torch.bfloat16)
    ref_cfg, cfg = parity.make_cfgs(tasks, svd_energy_threshold=0.9)
    # fp32 stand-in with identical deltas: base32 = base, ft32 = base + bf16(ft - base)
    base32 = {k: v.float() for k, v in base.items()}
    fts32 = {t: {k: base32[k] + (fts[t][k] - base[k]).float() for k in base} for t in tasks}
    ref = R.run_reference_path(base32, fts32, None, ref_cfg)
    res = merge_state_dicts(base, fts, None, cfg, "cuda", sign_ref={p: b["Vh"] for p, b in ref["bases"].items()})
    for name, b in ref["bases"].items():
        assert res["bases"].meta(name)["k"] == b["k"]
    for name in base:
        d_ref = ref["merged_deltas"][name]
        d_new = res["merged_state_dict"][name].cpu() - base32[name]
        assert res["merged_state_dict"][name].dtype == torch.float32
        assert parity.rel_l2(d_new, d_ref) < 2e-3, name


# ---- wide path: 17..32 task vectors (BASELINE configs[3]: 20 tasks, RTVQ sweep) -----------------------------
@pytest.mark.parametrize("n_tasks,bits,stages,strategy,mask_p", [
    (20, 4, 2, "union", 0.3), (20, 2, 4, "majority", 0.5), (20, 8, 1, "intersection", 0.9), (20, 4, 3, "union", None),
    (17, 4, 2, "majority", 0.5), (32, 4, 2, "union", 0.3)])
def test_wide_path_20_tasks_rtvq_sweep(cuda_device, n_tasks, bits, stages, strategy, mask_p):
    shapes = {"blk.attn.weight": (300, 70), "blk.bias": (4099,), "wide.weight": (1, 1, 40000), "ln.weight": (768,)}
    ref, res, _ = parity.run_both(shapes, n_tasks, mask_p=mask_p, svd_mask_strategy=strategy,
                                  svd_energy_threshold=0.9, svd_low_bits=bits, svd_rtvq_stages=stages)
    rep = parity.compare_run(ref, res)
    print(_summary(rep))
    assert rep["code_equal"] >= 0.98 * rep["code_total"]          # observed 0.9888 .. 1.0 (profiles/parity_rates.json)


def test_wide_path_cluster_and_fp32_basis(cuda_device):
    shapes = {"a.weight": (257, 64), "b": (5000,)}
    ref, res, _ = parity.run_both(shapes, 20, mask_p=0.9, svd_mask_strategy="intersection", svd_weighting="cluster",
                                  svd_cluster_k=2, svd_energy_threshold=0.9, svd_fp16=False)
    rep = parity.compare_run(ref, res)
    print(_summary(rep))


# ---- noise region (svd_include_noise): second basis over the rows outside the combined mask ----------------
@pytest.mark.parametrize("strategy,mask_p,shrink,weighting,n_tasks,fp16", [
    ("majority", 0.5, 0.5, "uniform", 8, True), ("union", 0.3, 0.25, "performance", 8, True),
    ("intersection", 0.9, 1.0, "cluster", 8, True), ("majority", 0.5, 0.5, "uniform", 5, False),
    ("union", 0.2, 0.5, "uniform", 12, True)])
def test_noise_region(cuda_device, strategy, mask_p, shrink, weighting, n_tasks, fp16):
    """svd_include_noise (cli.py:336-351, basis.py:455-466, compress.py:91-107, merge.py:257-284): the unmasked
    rows get their own basis, coefficients and codes, and shrink * reconstruction lands in the unmasked positions."""
    kw = dict(svd_weighting=weighting)
    if weighting == "cluster":
        kw["svd_cluster_k"] = 2
    if weighting == "performance":
        kw["svd_weighting_temperature"] = 5.0
    ref, res, _ = parity.run_both(parity.MEDIUM_SHAPES, n_tasks, mask_p=mask_p, svd_mask_strategy=strategy,
                                  svd_energy_threshold=0.9, svd_include_noise=True, svd_noise_shrink=shrink,
                                  svd_fp16=fp16, **kw)
    assert ref["bases_noise"], "the case must exercise the noise region"
    rep = parity.compare_run(ref, res)
    print(_summary(rep))
    assert rep["code_equal"] >= 0.99 * rep["code_total"]          # observed 0.9975 .. 1.0
    # unmasked positions are no longer zero deltas
    name = "blk.attn.weight"
    m = ref["combined_masks"][name]
    d = (res["merged_state_dict"][name].cpu() - ref["_base"][name])[~m]
    assert d.abs().max() > 0


def test_noise_region_edge_cases(cuda_device):
    """Noise region of a parameter whose mask is all-True (no noise basis), whose mask is below
    svd_min_mask_size (parameter skipped, no noise either), with a single unmasked element, and without a mask."""
    tasks = synth.task_names(4)
    shapes = {"alltrue": (64,), "small": (64,), "one_left": (64,), "nomask": (50,), "normal": (40, 40)}
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=5)
    masks = {}
    for i, t in enumerate(tasks):
        g = torch.Generator().manual_seed(100 + i)
        one = torch.ones(64, dtype=torch.bool)
        small = torch.zeros(64, dtype=torch.bool)
        small[:5] = True
        left = torch.ones(64, dtype=torch.bool)
        left[17] = False
        masks[t] = {"alltrue": one, "small": small, "one_left": left, "normal": torch.rand(40, 40, generator=g) < 0.5}
    ref_cfg, cfg = parity.make_cfgs(tasks, svd_mask_strategy="intersection", svd_energy_threshold=0.9,
                                    svd_include_noise=True, svd_noise_shrink=0.5)
    from oracle import svd_hybrid_ref as R
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    ref = R.run_reference_path(base, fts, masks, ref_cfg)
    ref["_base"] = base
    assert sorted(ref["bases_noise"]) == ["normal", "one_left"]
    res = merge_state_dicts(base, fts, masks, cfg, "cuda", sign_ref={p: b["Vh"] for p, b in ref["bases"].items()},
                            sign_ref_noise={p: b["Vh"] for p, b in ref["bases_noise"].items()})
    rep = parity.compare_run(ref, res)
    print(_summary(rep))
    assert torch.equal(res["merged_state_dict"]["small"].cpu(), base["small"])
    comp = res["compressed"]["alltrue"][tasks[0]]
    assert comp["unmasked"] is None and comp["masked"] is not None
    assert res["compressed"]["one_left"][tasks[0]]["unmasked"]["c_high_fp16"].numel() == 1


def test_noise_region_wide_path(cuda_device):
    shapes = {"blk.attn.weight": (300, 70), "blk.bias": (4099,), "ln.weight": (768,)}
    ref, res, _ = parity.run_both(shapes, 20, mask_p=0.5, svd_mask_strategy="majority", svd_energy_threshold=0.9,
                                  svd_include_noise=True, svd_noise_shrink=0.5)
    assert ref["bases_noise"]
    rep = parity.compare_run(ref, res)
    print(_summary(rep))
    assert rep["code_equal"] >= 0.99 * rep["code_total"]          # observed 1.0


def test_noise_region_materialised_bases(cuda_device):
    """bases[param]["noise"] in the reference layout (basis.py:455-466): U spans the same subspace."""
    ref, res, _ = parity.run_both({"w": (120, 50), "b": (3000,)}, 8, mask_p=0.5, svd_mask_strategy="majority",
                                  svd_energy_threshold=0.9, svd_include_noise=True)
    for name, rb in ref["bases_noise"].items():
        nb = res["bases"][name]["noise"]
        assert nb["U_high"].shape == rb["U_high"].shape and nb["U_low"].shape == rb["U_low"].shape
        assert nb["U_high"].dtype == rb["U_high"].dtype and nb["k"] == rb["k"] and nb["D"] == rb["D"]
        assert parity.max_principal_sine(nb["U_high"].float(), rb["U_high"].float()) < 2e-3   # fp16 storage
        assert torch.allclose(nb["mean"].cpu(), rb["mean"], atol=1e-7)
        mb, rmb = res["bases"][name]["masked"], ref["bases"][name]
        assert parity.max_principal_sine(mb["U_high"].float(), rmb["U_high"].float()) < 2e-3


def test_wide_path_materialised_bases(cuda_device):
    """U_high / U_low / mean in the reference layout for 20 task vectors (K5 with a runtime task count), incl.
    the noise basis: same subspaces and means as the oracle's LAPACK bases."""
    ref, res, _ = parity.run_both({"w": (130, 60), "b": (3001,)}, 20, mask_p=0.5, svd_mask_strategy="majority",
                                  svd_energy_threshold=0.9, svd_include_noise=True)
    for region, ref_bases in (("masked", ref["bases"]), ("noise", ref["bases_noise"])):
        assert ref_bases
        for name, rb in ref_bases.items():
            nb = res["bases"][name][region]
            assert nb["U_high"].shape == rb["U_high"].shape and nb["U_low"].shape == rb["U_low"].shape
            assert nb["U_high"].dtype == rb["U_high"].dtype and nb["k"] == rb["k"] and nb["D"] == rb["D"]
            assert parity.max_principal_sine(nb["U_high"].float(), rb["U_high"].float()) < 2e-3   # fp16 storage
            assert torch.allclose(nb["mean"].cpu(), rb["mean"], atol=1e-7)


@pytest.mark.parametrize("n_tasks", [1, 2, 3, 9, 16, 17])
@pytest.mark.parametrize("mask_p", [None, 0.5])
def test_task_count_sweep_with_tiny_parameters(cuda_device, n_tasks, mask_p):
    """Task counts at the seams of the kernel families (1, 2, 3; 9 and 16 on the direct kernels; 17 on the wide
    path) with a parameter of fewer elements than tasks (thin SVD with r = D < N) and one below svd_min_mask_size."""
    shapes = {"a.weight": (64, 33), "b": (700,), "c": (5,)}
    ref, res, _ = parity.run_both(shapes, n_tasks, mask_p=mask_p, svd_mask_strategy="majority",
                                  svd_energy_threshold=0.9)
    rep = parity.compare_run(ref, res)
    print(_summary(rep), rep["dust_params"])
    assert rep["chigh_equal"] >= 0.95 * rep["chigh_total"]        # observed 0.963 .. 1.0 (3 of 81 values one fp16 ulp off)


@pytest.mark.parametrize("dtype,n_tasks", [(torch.bfloat16, 20), (torch.float16, 20), (torch.bfloat16, 12),
                                           (torch.float16, 32)])
def test_half_precision_inputs_beyond_8_tasks(cuda_device, dtype, n_tasks):
    """bf16 / fp16 checkpoints through the direct (9..16 tasks) and wide (17..32 tasks) kernel families, with masks
    and a materialised basis; the oracle runs on the dtype-rounded deltas in fp32 (the reference's SVD rejects
    half-precision inputs on CPU)."""
    from oracle import svd_hybrid_ref as R
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    tasks = synth.task_names(n_tasks)
    shapes = {"w": (256, 130), "b": (4099,), "ln": (300,)}
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=21, dtype=dtype)
    masks = synth.make_masks(shapes, tasks, 0.5, seed=22)
    ref_cfg, cfg = parity.make_cfgs(tasks, svd_energy_threshold=0.9, svd_mask_strategy="majority")
    base32 = {k: v.float() for k, v in base.items()}
    fts32 = {t: {k: base32[k] + (fts[t][k] - base[k]).float() for k in base} for t in tasks}
    ref = R.run_reference_path(base32, fts32, masks, ref_cfg)
    res = merge_state_dicts(base, fts, masks, cfg, "cuda", sign_ref={p: b["Vh"] for p, b in ref["bases"].items()})
    for name, b in ref["bases"].items():
        assert res["bases"].meta(name)["k"] == b["k"], name
        nb = res["bases"][name]["masked"]
        assert nb["U_high"].shape == b["U_high"].shape
        assert parity.max_principal_sine(nb["U_high"].float(), b["U_high"].float()) < 2e-3
    for name in base:
        d_ref = ref["merged_deltas"][name]
        d_new = res["merged_state_dict"][name].cpu() - base32[name]
        assert res["merged_state_dict"][name].dtype == torch.float32
        assert parity.rel_l2(d_new, d_ref) < 2e-3, name


@pytest.mark.parametrize("kw", [dict(svd_max_rank=2), dict(svd_max_rank=None), dict(svd_min_mask_size=2000),
                                dict(svd_energy_threshold=1.0), dict(svd_energy_threshold=0.5),
                                dict(svd_center=False, svd_max_rank=3), dict(svd_low_bits=1, svd_rtvq_stages=4),
                                dict(svd_low_bits=8, svd_rtvq_stages=8)])
def test_config_corner_values(cuda_device, kw):
    """Rank cap, no cap, a mask-size gate that skips parameters, energy thresholds at the ends of (0, 1], and the
    extreme quantiser settings (select_rank basis.py:199-211, gate cli.py:332, rtvq.py:39-82)."""
    ref, res, _ = parity.run_both(parity.MEDIUM_SHAPES, 8, mask_p=0.5, svd_mask_strategy="majority", **kw)
    rep = parity.compare_run(ref, res)
    print(_summary(rep), rep["dust_params"])
    if "svd_max_rank" in kw and kw["svd_max_rank"] is not None:
        assert all(res["bases"].meta(n)["k"] <= kw["svd_max_rank"] for n in ref["bases"])
    if "svd_min_mask_size" in kw:
        assert set(ref["bases"]) < set(parity.MEDIUM_SHAPES)          # some parameters were skipped in both
