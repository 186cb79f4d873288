"""CPU tests of the host-side logic: config validation, task weighting, the Gram-based clustering,
the single-average equivalence behind cluster weighting, CLI parsing, artifact layout helpers,
and the parameter partition / collectives of the multi-GPU path (gloo, world size 2)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle import svd_hybrid_ref as R
from svd_quantization_task_merging_b200 import sharding, synth
from svd_quantization_task_merging_b200.svd_hybrid import basis, clustering, cli, storage, weighting
from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig


# ---- config -------------------------------------------------------------------------------------------
def test_config_defaults_and_validation():
    c = SVDHybridConfig()
    assert (c.svd_energy_threshold, c.svd_max_rank, c.svd_center, c.svd_fp16) == (0.95, 64, True, True)
    assert (c.svd_low_bits, c.svd_rtvq_stages, c.svd_mask_strategy, c.svd_weighting) == (4, 2, "union", "uniform")
    assert (c.svd_weighting_temperature, c.svd_cluster_k, c.svd_noise_shrink, c.svd_min_mask_size) == (5.0, 2, 0.5, 10)
    assert c.device == "cuda" and c.output_dir == "./svd_hybrid_output" and c.artifact_dir == "./artifacts"
    for bad in (dict(svd_mask_strategy="xor"), dict(svd_weighting="best"), dict(svd_energy_threshold=0.0),
                dict(svd_energy_threshold=1.5), dict(svd_low_bits=0), dict(svd_low_bits=9), dict(svd_rtvq_stages=0)):
        with pytest.raises(ValueError):
            SVDHybridConfig(**bad)


def test_config_field_order_matches_reference_config_json():
    """config.json is asdict(config): field names/order are part of the artifact layout."""
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_meta.json")))
    import dataclasses
    case = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pipeline_golden.pt"), weights_only=False)
    ref_keys = list(case["union_uniform"]["config_json"].keys())
    assert [f.name for f in dataclasses.fields(SVDHybridConfig)] == ref_keys
    assert gold["generator"].endswith("make_golden.py")


# ---- weights ------------------------------------------------------------------------------------------
def test_weights_match_oracle():
    tasks = synth.task_names(8)
    assert weighting.compute_weights(tasks, "uniform") == R.task_weights(R.RefConfig(tasks=tasks))
    perf = synth.performance_table(tasks)
    w = weighting.compute_performance_weights(perf, 5.0)
    assert w == R.task_weights(R.RefConfig(tasks=tasks, svd_weighting="performance", performance=perf))
    assert abs(sum(w.values()) - 1) < 1e-6 and w[tasks[-1]] > w[tasks[0]]
    assign = {t: i % 3 for i, t in enumerate(tasks)}
    assert weighting.compute_cluster_weights(tasks, assign) == R.task_weights(
        R.RefConfig(tasks=tasks, svd_weighting="cluster"), assign)
    with pytest.raises(ValueError):
        weighting.compute_weights(tasks, "nope")
    assert weighting.compute_weights(tasks, "performance", None) == weighting.compute_uniform_weights(tasks)


def test_performance_file_lookup(tmp_path):
    f = tmp_path / "perf.json"
    f.write_text(json.dumps({"Cars": 0.7, "euro_sat": 0.9}))
    m = weighting.load_performance_metrics(str(f), ["Cars", "EuroSAT", "DTD"])
    assert m == {"Cars": 0.7, "EuroSAT": 0.9, "DTD": 1.0}


def test_effective_weights_reproduce_clustered_merge():
    """merge_with_clustering (merge.py:555-626) == ONE weighted average with effective_merge_weights."""
    tasks = synth.task_names(6)
    shapes = {"w": (40, 30)}
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=5)
    assign = {t: (0 if i < 2 else 1) for i, t in enumerate(sorted(tasks))}
    cfg = R.RefConfig(tasks=tasks, svd_weighting="cluster", svd_energy_threshold=0.8)
    ref = R.run_reference_path(base, fts, None, cfg, assignments=assign)
    eff = weighting.effective_merge_weights(tasks, ref["weights"], assign)
    assert abs(sum(eff.values()) - 1.0) < 1e-6
    single = R.merge_deltas(ref["compressed"], ref["bases"], {}, eff, {"w": base["w"].shape})
    a, b = single["w"], ref["merged_deltas"]["w"]
    assert (a - b).norm() / b.norm() < 1e-6


# ---- clustering from the Gram ------------------------------------------------------------------------
def test_exact_kmeans_on_gram_matches_sklearn_on_clustered_tasks():
    rng = np.random.default_rng(0)
    for trial in range(8):
        N, P = 8, 400
        names = [f"t{(i * 3) % N}" for i in range(N)]
        lab = rng.integers(0, 2, N)
        lab[:2] = [0, 1]
        cen = rng.standard_normal((2, P))
        X = np.stack([cen[lab[i]] * 0.8 + rng.standard_normal(P) for i in range(N)])
        G = X @ X.T
        a = clustering.cluster_from_gram(G, names, 2, backend="exact")
        b = clustering.cluster_from_gram(G, names, 2, backend="sklearn")
        tv = {n: {"w": torch.from_numpy(X[i].astype(np.float32))} for i, n in enumerate(names)}
        c = R.cluster_tasks_full(tv, 2)
        for x in names:
            for y in names:
                assert (a[x] == a[y]) == (b[x] == b[y]) == (c[x] == c[y])
    with pytest.raises(ValueError):
        clustering.cluster_from_gram(np.eye(3), ["a", "b", "c"], 4)
    with pytest.raises(ValueError):
        clustering.cluster_from_gram(np.eye(3), ["a", "b", "c"], 2, method="dbscan", backend="sklearn")


def test_kmeans_inertia_is_global_optimum_small():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((7, 5))
    g = clustering.normalised_gram(X @ X.T)
    lab = clustering.kmeans_partition_from_gram(g, 3)
    rows = clustering._restricted_growth(7, 3, 10 ** 6)
    best = clustering._inertia(rows[rows.max(1) == 2], g, 3).min()
    assert abs(clustering._inertia(lab[None, :], g, 3)[0] - best) < 1e-12 and lab[0] == 0


# ---- rank selection (host mirror) ----------------------------------------------------------------------
def test_select_rank_host_mirror_matches_reference_golden():
    cases = torch.load(os.path.join(os.path.dirname(__file__), "golden", "rank_golden.pt"), weights_only=False)
    for c in cases:
        near_tie = bool(((c["cum"] - np.float32(c["thr"])).abs() < 2e-7).any())
        k = basis.select_rank(c["S"], c["thr"], c["max_rank"], c.get("min_rank", 1))
        assert k == c["k"] or near_tie
        assert (basis.compute_energy_spectrum(c["S"]) - c["cum"]).abs().max() <= 2e-7
    with pytest.raises(ValueError):
        basis.stack_and_center([])


# ---- CLI -----------------------------------------------------------------------------------------------
def test_cli_flags_and_json_overlays(tmp_path):
    a = cli.parse_args(["--tasks", "A", "B", "--checkpoint-dir", "c", "--base-model-path", "b.pt", "--no-fp16",
                        "--mask-strategy", "majority", "--weighting", "cluster", "--cluster-k", "3", "--rtvq-stages", "3",
                        "--store-artifacts", "--no-eval-reconstruction", "--energy-threshold", "0.9"])
    c = cli.config_from_args(a)
    assert c.tasks == ["A", "B"] and not c.svd_fp16 and c.svd_mask_strategy == "majority" and c.svd_cluster_k == 3
    assert c.svd_store_artifacts and not c.svd_eval_reconstruction and c.svd_energy_threshold == 0.9
    d = cli.config_from_args(cli.parse_args(["--tasks", "A", "--checkpoint-dir", "c", "--base-model-path", "b"]))
    assert (d.svd_energy_threshold, d.svd_max_rank, d.svd_weighting_temperature, d.svd_store_artifacts) == (0.95, 64, 5.0, False)
    q = tmp_path / "q.json"
    q.write_text(json.dumps({"tasks": ["X", "Y"], "checkpoints": {"checkpoint_dir": "cd", "base_model_path": "bp"},
                             "quantization": {"method": "asymmetric", "task_bits": 3}}))
    e = cli.config_from_args(cli.parse_args(["--quantize-config", str(q)]))
    assert e.tasks == ["X", "Y"] and e.checkpoint_dir == "cd" and e.base_model_path == "bp" and e.svd_low_bits == 4
    with pytest.raises(ValueError):
        cli.config_from_args(cli.parse_args(["--checkpoint-dir", "c", "--base-model-path", "b"]))


# ---- artifact layout --------------------------------------------------------------------------------------
def test_artifact_roundtrip_layout(tmp_path):
    basis_d = {"masked": {"U_high": torch.randn(10, 2).half(), "U_low": torch.randn(10, 3).half(),
                          "singular_values": torch.rand(5), "k": 2, "mean": torch.randn(10, 1), "energy_retained": 0.93,
                          "D": 10, "N": 5}, "noise": None}
    comp = {"a/b.weight": {"T1": {"masked": {"c_high_fp16": torch.randn(2).half(),
                                             "c_low_quant": {"payloads": [{"stage": 0, "quantized": torch.zeros(3, dtype=torch.uint8),
                                                                           "scale": torch.tensor(2.0), "zero_point": torch.tensor(1.0),
                                                                           "residual_norm": 0.5}],
                                                             "num_bits": 4, "num_stages": 1,
                                                             "original_shape": torch.Size([3]), "original_dtype": "torch.float32"}},
                                  "unmasked": None}}}
    diag = {"per_parameter": {"a/b.weight": {"original_shape": [2, 5], "masked_size": np.int64(10)}},
            "task_weights": {"T1": 1.0}, "summary": {"x": np.float32(1.5)}}
    cfg = SVDHybridConfig(tasks=["T1"], artifact_dir=str(tmp_path))
    storage.save_all_artifacts({"a/b.weight": basis_d}, comp, diag, cfg, str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["basis", "coeffs", "config.json", "diagnostics.json"]
    assert os.listdir(tmp_path / "basis") == ["a_b.weight.pt"] and os.listdir(tmp_path / "coeffs") == ["a_b.weight.pt"]
    art = storage.load_all_artifacts(str(tmp_path))
    assert art["config"] == cfg and art["diagnostics"]["per_parameter"]["a/b.weight"]["masked_size"] == 10
    assert list(art["bases"]["a/b.weight"].keys()) == ["masked"]
    assert list(art["bases"]["a/b.weight"]["masked"].keys()) == ["U_high", "U_low", "singular_values", "k", "mean",
                                                                 "energy_retained", "D", "N"]
    assert list(art["compressed"]["a/b.weight"]["T1"].keys()) == ["masked"]
    with pytest.raises(FileNotFoundError):
        storage.load_basis("missing", str(tmp_path))


# ---- multi-GPU host logic ------------------------------------------------------------------------------
def test_lpt_partition_balanced_and_deterministic():
    for model in ("ViT-B-32", "ViT-L-14", "Llama-3-8B"):
        shapes = synth.model_shapes(model)
        cost = {k: int(np.prod(v)) * 9 for k, v in shapes.items()}
        for world in (1, 2, 4, 8):
            owner = sharding.lpt_partition(cost, world)
            assert owner == sharding.lpt_partition(dict(reversed(list(cost.items()))), world)
            assert set(owner) == set(cost) and set(owner.values()) <= set(range(world))
            bal = sharding.partition_balance(cost, owner, world)
            assert bal < (1.08 if model != "Llama-3-8B" else 1.12), (model, world, bal)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = {"a": (3, 4), "b": (5,), "c": (2, 2, 2), "d": (7,)}
        cost = {k: int(np.prod(v)) for k, v in shapes.items()}
        owner = sharding.lpt_partition(cost, world)
        local = {n: torch.full(shapes[n], float(ord(n[0]) + 100 * r)) for n, r in owner.items() if r == rank}
        full = sharding.replicate_tensors(local, owner, shapes, torch.float32, "cpu")
        ok = all(torch.equal(full[n], torch.full(shapes[n], float(ord(n[0]) + 100 * owner[n]))) for n in shapes)
        g = torch.full((4, 4), float(rank + 1), dtype=torch.float64)
        sharding.allreduce_gram(g)
        ok = ok and bool((g == sum(range(1, world + 1))).all())
        recs = sharding.gather_objects({"rank": rank, "n": len(local)})
        ok = ok and [r["rank"] for r in recs] == list(range(world)) and sum(r["n"] for r in recs) == len(shapes)
        # flat record gather: ragged row counts (rank r contributes r + 1 rows), one collective
        rows = [q + 1 for q in range(world)]
        mine = np.full((rows[rank], 5), float(rank)) + np.arange(5)[None, :]
        got = sharding.gather_records(mine, rows, "cpu")
        ok = ok and len(got) == world and all(g.shape == (rows[q], 5) and (g == q + np.arange(5)[None, :]).all()
                                              for q, g in enumerate(got))
        # padded all-gather of the ranks' merged arenas
        sizes = [3 + 2 * q for q in range(world)]
        flat = torch.arange(sizes[rank], dtype=torch.float32) + 10 * rank
        allm = sharding.gather_merged(flat, sizes)
        ok = ok and all(torch.equal(allm[q, : sizes[q]], torch.arange(sizes[q], dtype=torch.float32) + 10 * q)
                        for q in range(world))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_collectives_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_safetensors_checkpoints_are_an_additive_input_format(tmp_path):
    """load_checkpoint (reference: task_vector_loader.py:56-100 reads pickles only) also accepts .safetensors files,
    and get_task_checkpoint_paths finds them after the reference's own candidates."""
    from safetensors.torch import save_file
    from svd_quantization_task_merging_b200.svd_hybrid.task_vector_loader import (get_task_checkpoint_paths,
                                                                                   load_checkpoint, load_task_vectors)
    g = torch.Generator().manual_seed(0)
    base = {"a.weight": torch.randn(4, 5, generator=g), "b": torch.randn(7, generator=g)}
    ft = {k: v + 0.01 * torch.randn(v.shape, generator=g) for k, v in base.items()}
    torch.save(base, tmp_path / "base.pt")
    os.makedirs(tmp_path / "ck")
    save_file(ft, str(tmp_path / "ck" / "A.safetensors"))
    torch.save(ft, tmp_path / "ck" / "B.pt")
    paths = get_task_checkpoint_paths(str(tmp_path / "ck"), ["A", "B"])
    assert paths["A"].endswith("A.safetensors") and paths["B"].endswith("B.pt")
    sd = load_checkpoint(paths["A"])
    assert sorted(sd) == sorted(ft) and all(torch.equal(sd[k], ft[k]) for k in ft)
    tv = load_task_vectors(str(tmp_path / "base.pt"), paths)
    assert all(torch.equal(tv["A"][k], tv["B"][k]) for k in base)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU arm: the oracle port on the box's host cores) prints ONE JSON line with the
    keys the driver reads, names the workload exactly as the GPU arm does, and needs no GPU."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "toy",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "params/s"
    assert d["metric"].startswith("task-vector params merged/sec")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert d["config"]["workload"] == b._workload_string("toy")


def test_entry_scripts_keep_the_reference_command_line():
    """scripts/reload_svd_hybrid.py and scripts/run_svd_hybrid.py expose the flags of the reference's scripts
    (scripts/reload_svd_hybrid.py:149-170; run_svd_hybrid.py forwards to the pipeline CLI)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "scripts", "reload_svd_hybrid.py")).read()
    flags = set(re.findall(r'add_argument\(\s*"(--[a-z-]+)"', text))
    assert flags == {"--artifact-dir", "--output-path", "--verify", "--merged-model-path", "--eval", "--eval-script",
                     "--verbose"}
    for helper in ("compute_state_dict_checksum", "verify_reconstruction", "run_evaluation", "main"):
        assert f"def {helper}(" in text
    run = open(os.path.join(root, "scripts", "run_svd_hybrid.py")).read()
    assert "from src.svd_hybrid.cli import main" in run
    from svd_quantization_task_merging_b200.svd_hybrid import storage
    assert callable(storage.load_merged_model)
