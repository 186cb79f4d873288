"""Size-independent properties at BASELINE.json's full sizes (the oracle needs minutes there, so the
full-size checks are properties the path must satisfy exactly):
  * determinism: two runs are bit-identical;
  * all-true masks == no masks, bit for bit;
  * power-of-two scale equivariance: ft' = base + 2*delta gives identical ranks, fp16/RTVQ codes
    and exactly doubled singular values / coefficients (every step of the path is homogeneous);
  * merged == base exactly where the combined mask is false; counts match the masks;
  * sharded (logical world size 8) == unsharded, bit for bit.
Plus engine edge cases: non-contiguous / misaligned views, mixed dtypes, integer buffers."""
import numpy as np
import pytest
import torch

from svd_quantization_task_merging_b200 import sharding, synth
from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig

pytestmark = pytest.mark.gpu


def _inputs(model, n_tasks, p, scale=1.0, seed=1234):
    shapes = synth.model_shapes(model)
    tasks = synth.task_names(n_tasks)
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=seed, device="cuda")
    if scale != 1.0:
        fts = {t: {k: base[k] + scale * (v - base[k]) for k, v in sd.items()} for t, sd in fts.items()}
    masks = synth.make_masks(shapes, tasks, p, seed=seed + 1, device="cuda") if p is not None else None
    return shapes, tasks, base, fts, masks


def test_vit_b_32_full_model_properties(cuda_device):
    """configs[0]: ViT-B-32 (87,849,216 params), 8 tasks, union masks, energy 0.9, 4-bit x 2 RTVQ, uniform."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    shapes, tasks, base, fts, masks = _inputs("ViT-B-32", 8, 0.3)
    assert synth.total_params(shapes) == 87_849_216
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy="union", svd_store_artifacts=False,
                          svd_eval_reconstruction=False)
    a = MergeJob(base, fts, masks, cfg, "cuda").run()
    b = MergeJob(base, fts, masks, cfg, "cuda").run()
    ma, mb = a.merged_state_dict(), b.merged_state_dict()
    fa = a._fetch()[torch.float32]
    assert all(torch.equal(ma[k], mb[k]) for k in ma)                                  # determinism
    assert (fa["info"][:, 0] == 0).all() and len(fa["info"]) == 152
    cm = a.combined_masks()
    union = {k: torch.stack([masks[t][k] for t in tasks]).any(0) for k in shapes}
    names = a.groups[torch.float32].names
    for i, k in enumerate(names):
        assert torch.equal(cm[k], union[k]), k                                        # masks bit-exact at full size
        assert int(fa["dm"][i]) == int(union[k].sum())
        assert torch.equal(ma[k][~union[k]], base[k][~union[k]]), k                   # untouched where mask is false
        assert torch.isfinite(ma[k]).all(), k
    # sharded == unsharded (logical world size 8)
    owner = sharding.lpt_partition({k: int(np.prod(v)) * 9 for k, v in shapes.items()}, 8)
    for rank in (0, 5):
        mine = [n for n, r in owner.items() if r == rank]
        part = MergeJob(base, fts, masks, cfg, "cuda", param_filter=mine).run().merged_state_dict()
        assert all(torch.equal(part[n], ma[n]) for n in mine)


def test_vit_b_32_twenty_tasks_wide_path_properties(cuda_device):
    """configs[3] style at full ViT-B-32 size: 20 tasks take the wide path (mask pack, staged single-pass Gram,
    runtime-N pass 2).  Determinism, bit-exact masks, untouched elements, sharded == unsharded, and the singular
    values of the largest parameters against an independent fp64 torch.linalg.svdvals of the masked, centred
    task matrix built with torch ops on the GPU."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    shapes, tasks, base, fts, masks = _inputs("ViT-B-32", 20, 0.3)
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_mask_strategy="majority",
                          svd_store_artifacts=False, svd_eval_reconstruction=False)
    a = MergeJob(base, fts, masks, cfg, "cuda").run()
    assert a.wide
    b = MergeJob(base, fts, masks, cfg, "cuda").run()
    ma, mb = a.merged_state_dict(), b.merged_state_dict()
    assert all(torch.equal(ma[k], mb[k]) for k in ma)                                  # determinism
    fa = a._fetch()[torch.float32]
    cm = a.combined_masks()
    names = a.groups[torch.float32].names
    big = sorted(range(len(names)), key=lambda i: -int(np.prod(shapes[names[i]])))[:3]
    for i, k in enumerate(names):
        votes = torch.stack([masks[t][k] for t in tasks]).sum(0)
        maj = 2 * votes >= len(tasks)
        assert torch.equal(cm[k], maj), k
        assert int(fa["dm"][i]) == int(maj.sum())
        assert torch.equal(ma[k][~maj], base[k][~maj]), k
        assert torch.isfinite(ma[k]).all(), k
        if i in big:
            T = torch.stack([(fts[t][k] - base[k])[maj] for t in tasks], 1).double()
            T = T - T.mean(1, keepdim=True)
            S = torch.linalg.svdvals(T).cpu().numpy()
            r = int(fa["info"][i, 2])
            assert r == 20
            got = fa["sv"][i][:r]
            ok = np.abs(got - S[:r]) <= np.maximum(2e-6 * S[0], 2e-9 * S[0] ** 2 / np.maximum(S[:r], 1e-30)) + \
                (S[:r] <= 1e-5 * S[0]) * 1e-4 * S[0]
            assert ok.all(), (k, got, S[:r])
    owner = sharding.lpt_partition({k: int(np.prod(v)) * 21 for k, v in shapes.items()}, 8)
    mine = [n for n, r in owner.items() if r == 3]
    part = MergeJob(base, fts, masks, cfg, "cuda", param_filter=mine).run().merged_state_dict()
    assert all(torch.equal(part[n], ma[n]) for n in mine)


def test_llama_3_8b_shard_bf16_capacity(cuda_device):
    """configs[4]: Llama-3-8B-shaped bf16 task vectors, parameter-sharded 8 ways.  Rank 0's shard of the LPT
    partition (lm_head: one 525,336,576-element parameter, 1.0 B parameters in all) on one GPU: the partition is
    balanced without row-splitting, the run is deterministic and finite, the spectrum of the 525 M x 8 task matrix
    matches an independent fp64 Gram eigensolve, and with uniform weights the merged delta is the mean task vector
    up to the coefficient quantisation (the centred coefficients average to zero)."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~80 GB of HBM")
    shapes_all = synth.model_shapes("Llama-3-8B")
    assert synth.total_params(shapes_all) == 8_030_261_248
    cost = {k: int(np.prod(v)) * 9 for k, v in shapes_all.items()}
    owner = sharding.lpt_partition(cost, 8)
    assert sharding.partition_balance(cost, owner, 8) < 1.01
    shapes = {k: v for k, v in shapes_all.items() if owner[k] == 0}
    assert "lm_head.weight" in shapes and 0.99e9 < synth.total_params(shapes) < 1.02e9
    tasks = synth.task_names(8)
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=77, device="cuda", dtype=torch.bfloat16)
    torch.cuda.empty_cache()
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_store_artifacts=False, svd_eval_reconstruction=False)
    a = MergeJob(base, fts, None, cfg, "cuda", diagnostics=False).run()
    ma = {k: v.clone() for k, v in a.merged_state_dict().items()}
    fa = a._fetch()[torch.bfloat16]
    names = a.groups[torch.bfloat16].names
    del a
    torch.cuda.empty_cache()
    b = MergeJob(base, fts, None, cfg, "cuda", diagnostics=False).run()
    mb = b.merged_state_dict()
    assert all(torch.equal(ma[k].view(torch.int32), mb[k].view(torch.int32)) for k in ma)      # determinism
    del b, mb
    torch.cuda.empty_cache()
    assert (fa["info"][:, 0] == 0).all() and len(names) == len(shapes)
    i = names.index("lm_head.weight")
    # independent spectrum: bf16 task vectors (ft - base rounded to bf16, task_vector_loader.py:103-145), centred
    # in fp64, Gram accumulated in fp64 over row chunks
    k = "lm_head.weight"
    G = torch.zeros(8, 8, dtype=torch.float64, device="cuda")
    mean_delta_sq = 0.0
    err_sq = 0.0
    flat_m, flat_b = ma[k].view(-1), base[k].view(-1)
    step = 1 << 24
    for lo in range(0, flat_b.numel(), step):
        T = torch.stack([(fts[t][k].view(-1)[lo: lo + step] - flat_b[lo: lo + step]) for t in tasks], 1).double()
        mean = T.mean(1)
        Tc = T - mean[:, None]
        G += Tc.T @ Tc
        got = flat_m[lo: lo + step].double() - flat_b[lo: lo + step].double()
        mean_delta_sq += float((mean ** 2).sum())
        err_sq += float(((got - mean) ** 2).sum())
        assert torch.isfinite(got).all()
    S = torch.linalg.eigvalsh(G).clamp_min(0).sqrt().flip(0).cpu().numpy()
    r = int(fa["info"][i, 2])
    assert r == 8
    got_s = fa["sv"][i][:r]
    ok = np.abs(got_s - S[:r]) <= np.maximum(2e-6 * S[0], 2e-9 * S[0] ** 2 / np.maximum(S[:r], 1e-30)) + \
        (S[:r] <= 1e-5 * S[0]) * 1e-4 * S[0]
    assert ok.all(), (got_s, S[:r])
    # merged - base = mean task vector + U cbar, and cbar ~ 0: the deviation is quantisation noise of the coefficients
    assert err_sq ** 0.5 <= 0.05 * (float(np.sum(S ** 2)) / 8) ** 0.5 + 1e-3 * mean_delta_sq ** 0.5, (err_sq, mean_delta_sq)


def test_all_true_masks_equal_no_masks_and_scale_equivariance(cuda_device):
    """ViT-B-16 shapes (configs[1] sizes), 8 tasks, 4-bit x 3 stages."""
    from svd_quantization_task_merging_b200.engine import MergeJob
    shapes, tasks, base, fts, _ = _inputs("ViT-B-16", 8, None)
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_rtvq_stages=3, svd_store_artifacts=False,
                          svd_eval_reconstruction=False)
    plain = MergeJob(base, fts, None, cfg, "cuda").run()
    ones = {t: {k: torch.ones(v, dtype=torch.bool, device="cuda") for k, v in shapes.items()} for t in tasks}
    masked = MergeJob(base, fts, ones, cfg, "cuda").run()
    mp, mm = plain.merged_state_dict(), masked.merged_state_dict()
    assert all(torch.equal(mp[k], mm[k]) for k in mp)
    # scale equivariance with an exact power of two.  Inputs snapped to a dyadic grid (base: multiples of
    # 2^-12, deltas: multiples of 2^-20) so that base + delta and base + 2 delta are exactly representable.
    qb = {k: torch.round(v * 4096.0) / 4096.0 for k, v in base.items()}
    qd = {t: {k: torch.round((v - base[k]) * 1048576.0) / 1048576.0 for k, v in sd.items()} for t, sd in fts.items()}
    f1x = {t: {k: qb[k] + d for k, d in sd.items()} for t, sd in qd.items()}
    f2x = {t: {k: qb[k] + 2.0 * d for k, d in sd.items()} for t, sd in qd.items()}
    assert all(torch.equal(f2x[t][k] - qb[k], 2.0 * (f1x[t][k] - qb[k])) for t in tasks for k in shapes)
    one = MergeJob(qb, f1x, None, cfg, "cuda").run()
    two = MergeJob(qb, f2x, None, cfg, "cuda").run()
    f1, f2 = one._fetch()[torch.float32], two._fetch()[torch.float32]
    assert np.array_equal(f1["info"], f2["info"])                                      # same ranks everywhere
    assert np.array_equal(f1["codes"], f2["codes"])                                    # identical RTVQ codes
    assert np.array_equal(2.0 * f1["sv"], f2["sv"])                                    # exactly doubled spectrum
    assert np.array_equal(2.0 * f1["coef"], f2["coef"])                                # exactly doubled coefficients
    assert np.array_equal(f1["qscale"], 2.0 * f2["qscale"]) and np.array_equal(f1["qzp"], f2["qzp"])
    # the fp16 high block is equivariant wherever fp16 is in its normal range (|c| >= 2^-14); below that
    # the fp16 grid itself is not scale-invariant (absolute spacing 2^-24)
    h1 = torch.from_numpy(f1["chigh"].copy()).view(torch.float16).float()
    h2 = torch.from_numpy(f2["chigh"].copy()).view(torch.float16).float()
    normal = h1.abs() >= 2.0 ** -14
    assert torch.equal(2.0 * h1[normal], h2[normal]) and (2.0 * h1 - h2).abs().max() <= 2.0 ** -23
    assert np.abs(2.0 * f1["cbar"] - f2["cbar"]).max() <= 2.0 ** -23
    m1, m2 = one.merged_state_dict(), two.merged_state_dict()
    for k in list(shapes)[:40]:
        d1, d2 = (m1[k].double() - qb[k].double()), (m2[k].double() - qb[k].double())
        assert (2.0 * d1 - d2).abs().max() <= 4e-7 * qb[k].abs().max().item() + 1e-9, k   # one rounding of base + delta


def test_engine_edge_cases_views_dtypes_and_buffers(cuda_device):
    from oracle import svd_hybrid_ref as R
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    tasks = synth.task_names(4)
    g = torch.Generator().manual_seed(5)
    big = torch.randn(4, 130, 67, generator=g)
    base = {"a.weight": big[1],                                   # view with a misaligned storage offset
            "b.weight": torch.randn(96, 40, generator=g).t(),      # non-contiguous
            "c.half": (torch.randn(300, generator=g) * 0.02).half(),
            "d.bf16": (torch.randn(64, 33, generator=g) * 0.02).bfloat16(),
            "steps": torch.tensor(12, dtype=torch.int64), "empty": torch.zeros(0)}
    q, _ = torch.linalg.qr(torch.randn(4, 4, generator=g))
    fts = {}
    for i, t in enumerate(tasks):
        sd = {}
        for k, v in base.items():
            if v.is_floating_point() and v.numel():
                mix = sum((q[i, j] * 0.6 ** j) * torch.randn(v.shape, generator=torch.Generator().manual_seed(100 + j)) * 0.01
                          for j in range(4))
                sd[k] = (v.float() + mix).to(v.dtype)
            else:
                sd[k] = v.clone()
        fts[t] = sd
    cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.8, svd_store_artifacts=False)
    res = merge_state_dicts(base, fts, None, cfg, "cuda")
    m = res["merged_state_dict"]
    assert list(m.keys()) == list(base.keys())
    assert torch.equal(m["steps"].cpu(), base["steps"]) and m["empty"].numel() == 0
    assert m["a.weight"].shape == base["a.weight"].shape and m["b.weight"].shape == base["b.weight"].shape
    assert m["c.half"].dtype == torch.float32 and m["d.bf16"].dtype == torch.float32        # base + fp32 delta promotes
    # fp32 tensors against the oracle (k and merged at RTVQ-noise level without sign hints)
    f32 = {k: v.contiguous() for k, v in base.items() if v.dtype == torch.float32 and v.numel()}
    ref = R.run_reference_path(f32, {t: {k: fts[t][k].contiguous() for k in f32} for t in tasks}, None,
                               R.RefConfig(tasks=tasks, svd_energy_threshold=0.8))
    for k in f32:
        assert res["bases"].meta(k)["k"] == ref["bases"][k]["k"]
        d_ref, d_new = ref["merged_deltas"][k], m[k].cpu() - f32[k]
        if torch.isfinite(d_ref).all():
            assert ((d_new - d_ref).norm() / d_ref.norm()).item() < 5e-2, k
