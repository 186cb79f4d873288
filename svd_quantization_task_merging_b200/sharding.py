"""Multi-GPU execution: one process per GPU, parameters partitioned across ranks.

Every parameter tensor is an independent unit of work from task-vector construction to the merged
weight (reference: the sequential loops of src/svd_hybrid/cli.py:317, compress.py:189, merge.py:371),
so the path shards with NO data-path collective.  NCCL (torch.distributed) is used only
  * to all-reduce the N x N whole-model task Gram (fp64) when cluster weighting needs it,
  * to gather the per-parameter records / diagnostics, and
  * optionally to replicate the merged tensors (one broadcast of a flat buffer per owner).
Results are bit-identical for any world size: the reduction order inside a parameter is fixed by
the tile size, never by the placement.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Mapping, Optional, Sequence

import torch
import torch.distributed as dist


def lpt_partition(cost: Mapping[str, int], world_size: int) -> Dict[str, int]:
    """Greedy longest-processing-time bin packing: parameter -> owning rank.  Deterministic
    (ties broken by name), so every rank computes the same map without communicating."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    load = [0] * world_size
    owner: Dict[str, int] = {}
    for name in sorted(cost, key=lambda n: (-int(cost[n]), n)):
        r = min(range(world_size), key=lambda i: (load[i], i))
        owner[name] = r
        load[r] += int(cost[name])
    return owner


def partition_balance(cost: Mapping[str, int], owner: Mapping[str, int], world_size: int) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    load = [0] * world_size
    for n, r in owner.items():
        load[r] += int(cost[n])
    mean = sum(load) / world_size
    return max(load) / mean if mean > 0 else 1.0


def allreduce_gram(gram: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank whole-model task Grams (N x N fp64).  Identity without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(gram, op=dist.ReduceOp.SUM, group=group)
    return gram


def gather_objects(obj, group=None) -> List:
    """Small python records (diagnostics, ranks, codes) from every rank, in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [obj]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def replicate_tensors(local: Mapping[str, torch.Tensor], owner: Mapping[str, int], shapes: Mapping[str, Sequence[int]],
                      dtype: torch.dtype, device, group=None) -> "OrderedDict[str, torch.Tensor]":
    """Give every rank every owner's tensors: per owner ONE broadcast of a flat buffer holding its
    tensors in sorted-name order.  ``local`` holds the tensors this rank owns."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for r in range(world):
        names = sorted(n for n, o in owner.items() if o == r)
        if not names:
            continue
        sizes = [int(torch.Size(shapes[n]).numel()) for n in names]
        flat = torch.empty(sum(sizes), dtype=dtype, device=device)
        if r == rank:
            off = 0
            for n, s in zip(names, sizes):
                flat[off: off + s].copy_(local[n].reshape(-1))
                off += s
        if world > 1:
            dist.broadcast(flat, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        off = 0
        for n, s in zip(names, sizes):
            out[n] = flat[off: off + s].view(torch.Size(shapes[n]))
            off += s
    return out


def merge_state_dicts_sharded(base, finetuned, task_masks, config, device: Optional[str] = None, group=None,
                              replicate_merged: bool = False, **kw) -> Dict:
    """Parameter-sharded SVD-Hybrid merge.  Every rank passes the same (or at least its own shard of
    the) inputs; rank r processes the parameters ``lpt_partition`` assigns to it.

    Returns this rank's result dict (merged tensors of its shard; diagnostics and per-parameter
    records gathered from all ranks on every rank; all merged tensors when ``replicate_merged``)."""
    from .engine import MergeJob
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n_tasks = len(config.tasks) if config.tasks else len(finetuned)
    cost = {k: int(v.numel()) * (n_tasks + 1) for k, v in base.items() if torch.is_tensor(v) and v.is_floating_point()}
    owner = lpt_partition(cost, world)
    mine = [n for n, r in owner.items() if r == rank]
    job = MergeJob(base, finetuned, task_masks, config, device, param_filter=mine, **kw)
    if world > 1 and job.cluster_mode:
        job.gram_reduce_hook = lambda g: allreduce_gram(g, group)
    job.run()
    res = job.results()
    local_merged = {n: res["merged_state_dict"][n] for n in mine if n in res["merged_state_dict"]}
    records = {"rank": rank, "params": {n: res["bases"].meta(n) for n in res["bases"]},
               "per_parameter": res["diagnostics"].get("per_parameter", {})}
    gathered = gather_objects(records, group)
    per_parameter, meta = {}, {}
    for rec in gathered:
        per_parameter.update(rec["per_parameter"])
        meta.update(rec["params"])
    diag = dict(res["diagnostics"])
    if per_parameter:
        from .svd_hybrid.diagnostics import summarize
        diag["per_parameter"] = OrderedDict(sorted(per_parameter.items()))
        diag["summary"] = summarize(diag["per_parameter"])
    out = {"merged_state_dict": local_merged, "diagnostics": diag, "bases": res["bases"],
           "compressed": res["compressed"], "owner": owner, "basis_meta": meta, "job": job}
    if replicate_merged:
        shapes = {n: tuple(base[n].shape) for n in owner}
        full = replicate_tensors({n: t.float() for n, t in local_merged.items()}, owner, shapes, torch.float32,
                                 job.device, group)
        merged_all = OrderedDict()
        for k, v in base.items():
            merged_all[k] = full[k] if k in full else (v.clone() if torch.is_tensor(v) else v)
        out["merged_state_dict"] = merged_all
    return out
