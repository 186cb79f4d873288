"""Multi-GPU execution: one process per GPU, parameters partitioned across ranks.

Every parameter tensor is an independent unit of work from task-vector construction to the merged
weight (reference: the sequential loops of src/svd_hybrid/cli.py:317, compress.py:189, merge.py:371),
so the path shards with NO data-path collective.  NCCL (torch.distributed) is used only
  * to all-reduce the N x N whole-model task Gram (fp64) when cluster weighting needs it,
  * to gather the per-parameter records / diagnostics (one flat fp64 all_gather_into_tensor), and
  * optionally to replicate the merged tensors (one broadcast of a flat buffer per owner, or one padded
    all_gather_into_tensor of the ranks' output arenas).
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


def gather_records(local: "np.ndarray", rows_per_rank: Sequence[int], device, group=None) -> List["np.ndarray"]:
    """Per-parameter fp64 records of every rank with ONE collective: each rank contributes a [rows_per_rank[r], W]
    matrix, padded to the longest, through all_gather_into_tensor.  -> list of [rows_r, W] arrays in rank order."""
    import numpy as np
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return [np.asarray(local, np.float64)]
    width = int(local.shape[1])
    pmax = max(int(n) for n in rows_per_rank)
    buf = torch.zeros(max(pmax, 1), width, dtype=torch.float64, device=device)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, np.float64)).to(device)
    out = torch.empty(world * max(pmax, 1), width, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf, group=group)
    host = out.view(world, max(pmax, 1), width).cpu().numpy()
    return [host[r, : int(rows_per_rank[r])] for r in range(world)]


def gather_merged(local_flat: torch.Tensor, sizes_per_rank: Sequence[int], group=None) -> torch.Tensor:
    """All ranks' merged shards with ONE padded all_gather_into_tensor over NVLink: rank r contributes its flat fp32
    arena (sizes_per_rank[r] elements).  -> [world, max size] tensor; row r's first sizes_per_rank[r] elements are
    rank r's arena."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    smax = max(int(x) for x in sizes_per_rank)
    if world == 1:
        return local_flat.view(1, -1)
    buf = local_flat
    if buf.numel() != smax:
        buf = torch.zeros(smax, dtype=local_flat.dtype, device=local_flat.device)
        buf[: local_flat.numel()] = local_flat
    out = torch.empty(world * smax, dtype=local_flat.dtype, device=local_flat.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out.view(world, smax)


def merge_state_dicts_sharded(base, finetuned, task_masks, config, device: Optional[str] = None, group=None,
                              replicate_merged: bool = False, owner: Optional[Mapping[str, int]] = None,
                              to_host=False, **kw) -> Dict:
    """Parameter-sharded SVD-Hybrid merge.  Every rank passes the same inputs, or just its own shard of them
    together with the global ``owner`` map; rank r processes the parameters ``lpt_partition`` assigns to it and
    uploads only those (engine: the filter is applied before any host->device copy).

    Collectives: the N x N fp64 whole-model Gram all-reduce of cluster weighting (on the job's side stream), ONE
    flat all-gather of the per-parameter records / diagnostics, and -- only with ``replicate_merged`` -- one
    broadcast of a flat buffer per owner.  Returns this rank's result dict (merged tensors of its shard;
    diagnostics gathered from all ranks on every rank; all merged tensors when ``replicate_merged``)."""
    from .engine import MergeJob
    from .results import diagnostics_from_records, pack_records, record_width
    import numpy as np
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n_tasks = len(config.tasks) if config.tasks else len(finetuned)
    if owner is None:
        cost = {k: int(v.numel()) * (n_tasks + 1) for k, v in base.items()
                if torch.is_tensor(v) and v.is_floating_point()}
        owner = lpt_partition(cost, world)
    by_rank = [sorted(n for n, r in owner.items() if r == q) for q in range(world)]
    mine = by_rank[rank]
    import time as _time
    job = MergeJob(base, finetuned, task_masks, config, device, param_filter=mine, **kw)
    if world > 1 and job.cluster_mode:
        job.gram_reduce_hook = lambda g: allreduce_gram(g, group)
    t0 = _time.perf_counter()
    job.run()
    if job._profile:
        torch.cuda.synchronize(job.device)
        job.timing["run_s"] = _time.perf_counter() - t0
        t0 = _time.perf_counter()
    res = job.results(to_host=to_host)
    if job._profile:
        torch.cuda.synchronize(job.device)
        job.timing["results_s"] = _time.perf_counter() - t0
        t0 = _time.perf_counter()
    local_merged = {n: res["merged_state_dict"][n] for n in mine if n in res["merged_state_dict"]}
    # one fp64 record per owned parameter, rows in sorted-name order (status -1: no basis), gathered flat
    index = res["bases"]._index
    names_l, rec_l = pack_records(job, index)
    width = record_width(job.N)
    local = np.full((len(mine), width), -1.0, np.float64)
    pos = {n: i for i, n in enumerate(mine)}
    for n, row in zip(names_l, rec_l):
        local[pos[n]] = row
    gathered = gather_records(local, [len(b) for b in by_rank], job.device, group)
    names_all, rows = [], []
    for q in range(world):
        for n, row in zip(by_rank[q], gathered[q]):
            if row[0] >= 0:
                names_all.append(n)
                rows.append(row)
    rec_all = np.stack(rows) if rows else np.zeros((0, width))
    meta = {n: {"k": int(r[3]), "r": int(r[2]), "D": int(r[4]), "N": int(r[1]), "energy_retained": float(r[8]),
                "solved": int(r[0]) == 0} for n, r in zip(names_all, rec_all)}
    diag = dict(res["diagnostics"])
    if job.want_diag:
        full = diagnostics_from_records(job.cfg, job.tasks, job.bits, job.stages, names_all, rec_all)
        diag["per_parameter"] = full["per_parameter"]
        diag["summary"] = full["summary"]
    if job._profile:
        job.timing["gather_records_s"] = _time.perf_counter() - t0
    out = {"merged_state_dict": local_merged, "diagnostics": diag, "bases": res["bases"],
           "compressed": res["compressed"], "owner": dict(owner), "basis_meta": meta, "job": job}
    if replicate_merged:
        shapes = {n: tuple(base[n].shape) for n in owner}
        full = replicate_tensors({n: t.float() for n, t in local_merged.items()}, owner, shapes, torch.float32,
                                 job.device, group)
        merged_all = OrderedDict()
        for k, v in base.items():
            merged_all[k] = full[k] if k in full else (v.clone() if torch.is_tensor(v) else v)
        out["merged_state_dict"] = merged_all
    return out
