"""Two-GPU check of the parameter-sharded path over NCCL (skipped on a single-GPU box): every rank
processes its LPT shard, the whole-model Gram is all-reduced for cluster weighting, merged tensors are
replicated with one broadcast per owner -- and the result equals the single-GPU run bit for bit."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, weighting, q):
    import torch.distributed as dist
    from svd_quantization_task_merging_b200 import sharding, synth
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig
    from tests import parity
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        tasks = synth.task_names(8)
        shapes = dict(parity.MEDIUM_SHAPES)
        shapes.update({f"extra{i}.weight": (64 + i, 33) for i in range(9)})
        base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=31)
        masks = synth.make_masks(shapes, tasks, 0.6, seed=32)
        cfg = SVDHybridConfig(tasks=tasks, svd_energy_threshold=0.9, svd_weighting=weighting, svd_store_artifacts=False)
        res = sharding.merge_state_dicts_sharded(base, fts, masks, cfg, f"cuda:{rank}", replicate_merged=True)
        ok = True
        if rank == 0:
            full = merge_state_dicts(base, fts, masks, cfg, "cuda:0")
            for k, v in full["merged_state_dict"].items():
                ok = ok and torch.equal(res["merged_state_dict"][k].cpu(), v.cpu())
            ok = ok and res["diagnostics"]["per_parameter"].keys() == full["diagnostics"]["per_parameter"].keys()
            for k, d in full["diagnostics"]["per_parameter"].items():
                ok = ok and res["diagnostics"]["per_parameter"][k] == d
            ok = ok and res["diagnostics"]["summary"] == full["diagnostics"]["summary"]
            if weighting == "cluster":
                ok = ok and res["job"].cluster_assignments == full["job"].cluster_assignments
        q.put((rank, bool(ok), len(res["merged_state_dict"])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("weighting", ["uniform", "cluster"])
def test_two_gpu_sharded_merge_equals_single_gpu(weighting):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, weighting, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[:2] for r in res] == [(0, True), (1, True)], res
