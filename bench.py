#!/usr/bin/env python
"""Headline benchmark: task-vector params merged / second through the SVD-Hybrid hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the whole hot path (K1 task vectors + tall masks + Gram, K2 per-parameter
solve incl. rank selection and 4-bit multi-stage RTVQ, K3 weighted reconstruction + merge) over one
synthetic checkpoint set.  Default workload (BASELINE.json configs[2], the configuration the metric
and the >= 60 % roofline target are quoted on): CLIP ViT-L-14 image encoder, 8 task vectors,
intersection tall masks, cluster weighting (k = 2), energy 0.9, 4-bit x 2-stage RTVQ, fp16 bases.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "task-vector params merged/sec (SVD+RTVQ+recon)"
UNIT = "params/s"

WORKLOADS = {
    # name: (model, n_tasks, mask strategy, mask p, weighting, stages)
    "vit-l-14-cluster": ("ViT-L-14", 8, "intersection", 0.9, "cluster", 2),
    "vit-b-16-majority": ("ViT-B-16", 8, "majority", 0.5, "performance", 3),
    "vit-b-32-union": ("ViT-B-32", 8, "union", 0.3, "uniform", 2),
    "toy": ("toy", 8, "union", 0.5, "uniform", 2),
    # configs[3]-style stress (not a bench line): 20 task vectors take the blocked-Gram wide path
    "vit-l-14-20tasks": ("ViT-L-14", 20, "union", 0.3, "uniform", 2),
    "vit-l-14-14tasks": ("ViT-L-14", 14, "majority", 0.5, "uniform", 2),
    # configs[4] capacity stress: one rank's shard of the 8-way LPT partition of Llama-3-8B (bf16 task vectors, no
    # masks); under torchrun with 8 ranks every rank takes its own shard = the whole model
    "llama-3-8b-shard": ("Llama-3-8B", 8, "union", None, "uniform", 2),
}
# workload extras: input dtype, logical world size of the parameter partition the shard is taken from
WORKLOAD_EXTRA = {"llama-3-8b-shard": {"dtype": "bfloat16", "shard_world": 8}}


def _measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples that arrived inside [t_begin, t_end] (the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if t_begin is not None and not (t_begin <= ts <= t_end + 0.03):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _make_cfg(workload):
    from svd_quantization_task_merging_b200 import synth
    from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig
    model, n_tasks, strategy, p, weighting, stages = WORKLOADS[workload]
    tasks = synth.task_names(n_tasks)
    cfg = SVDHybridConfig(tasks=tasks, model=model, svd_energy_threshold=0.9, svd_max_rank=64, svd_center=True,
                          svd_fp16=True, svd_low_bits=4, svd_rtvq_stages=stages, svd_mask_strategy=strategy,
                          svd_weighting=weighting, svd_weighting_temperature=5.0, svd_cluster_k=2,
                          svd_store_artifacts=False, svd_eval_reconstruction=False)
    return cfg, tasks, model, p


def _workload_string(workload: str, family: str = "parity") -> str:
    """The workload as both arms name it in config.workload (what is merged; not where or on what sample)."""
    import torch
    from svd_quantization_task_merging_b200 import synth
    cfg, tasks, model, p = _make_cfg(workload)
    shapes_all = synth.model_shapes(model)
    in_dtype = WORKLOAD_EXTRA.get(workload, {}).get("dtype", "float32")
    return (f"{workload}: {model} {'image encoder' if model.startswith('ViT') else 'weights'} "
            f"({len(shapes_all)} tensors, {synth.total_params(shapes_all)} params, {in_dtype}) x {len(tasks)} random-init "
            f"task vectors ({'decaying spectrum' if family == 'parity' else 'iid'}), "
            + (f"{cfg.svd_mask_strategy} tall masks (Bernoulli {p}), " if p is not None else "no masks, ")
            + f"{cfg.svd_weighting} weighting, energy {cfg.svd_energy_threshold}, {cfg.svd_low_bits}-bit x "
            f"{cfg.svd_rtvq_stages}-stage RTVQ, fp16 bases")


def _cpu_sample_names(shapes, model):
    """Bounded CPU sample: the tensors of the first four transformer blocks (the whole toy model)."""
    if model == "toy":
        return list(shapes.keys())
    prefs = tuple(f"transformer.resblocks.{i}." for i in range(4))
    names = [k for k in shapes if k.startswith(prefs)]
    if names:
        return names
    # other models: the attention projections and norms of the first decoder layer (Llama-3-8B: 41.9 M params)
    return [k for k in shapes if k.startswith("model.layers.0.") and ".mlp." not in k]


def cpu_baseline(workload: str, steps: int = 1, warmup: int = 0):
    """The reference's CPU path (oracle port: torch-eager ops, LAPACK SVD per parameter, per-task
    projections, CPU RTVQ) timed on this box's host cores on a bounded sample of the workload."""
    import torch
    from oracle import svd_hybrid_ref as R
    from svd_quantization_task_merging_b200 import synth
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        n_host = len(os.sched_getaffinity(0))
    except AttributeError:
        n_host = os.cpu_count() or 1
    if torch.get_num_threads() < n_host:
        torch.set_num_threads(n_host)
    cfg, tasks, model, p = _make_cfg(workload)
    shapes_all = synth.model_shapes(model)
    names = _cpu_sample_names(shapes_all, model)
    shapes = {k: shapes_all[k] for k in names}
    in_dtype = getattr(torch, WORKLOAD_EXTRA.get(workload, {}).get("dtype", "float32"))
    base, fts = synth.make_checkpoints(shapes, tasks, family="parity", seed=1234, dtype=in_dtype)
    masks = synth.make_masks(shapes, tasks, p, seed=4321) if p is not None else None
    upcast = ""
    if in_dtype != torch.float32:
        # torch.linalg.svd has no 16-bit CPU kernel, so the reference cannot run these inputs as they are: it gets
        # the fp32 stand-in with identical task vectors, base32 = base, ft32 = base + bf16(ft - base)
        fts = {t: {k: base[k].float() + (v - base[k]).float() for k, v in sd.items()} for t, sd in fts.items()}
        base = {k: v.float() for k, v in base.items()}
        upcast = f", {str(in_dtype).split('.')[-1]} task vectors upcast to fp32"
    perf = synth.performance_table(tasks) if cfg.svd_weighting == "performance" else None
    rcfg = R.RefConfig(tasks=tasks, svd_energy_threshold=cfg.svd_energy_threshold, svd_max_rank=cfg.svd_max_rank,
                       svd_center=cfg.svd_center, svd_fp16=cfg.svd_fp16, svd_low_bits=cfg.svd_low_bits,
                       svd_rtvq_stages=cfg.svd_rtvq_stages, svd_mask_strategy=cfg.svd_mask_strategy,
                       svd_weighting=cfg.svd_weighting, svd_weighting_temperature=cfg.svd_weighting_temperature,
                       svd_cluster_k=cfg.svd_cluster_k, svd_eval_reconstruction=False, performance=perf)
    # the reference's own clustering is k-means over the flattened [N x P_total] matrix (742 s of an
    # 863 s ViT-L-14 run); it is excluded from the bounded sample by injecting a
    # fixed partition, i.e. the CPU number is the reference's compute stages WITHOUT its slowest step.
    assign = {t: i % 2 for i, t in enumerate(sorted(tasks))} if cfg.svd_weighting == "cluster" else None
    n_params = synth.total_params(shapes)
    for _ in range(warmup):
        R.run_reference_path(base, fts, masks, rcfg, assignments=assign)
    t0 = time.perf_counter()
    for _ in range(max(steps, 1)):
        R.run_reference_path(base, fts, masks, rcfg, assignments=assign)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": n_params / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{model} {len(names)} tensors ({n_params / 1e6:.1f} M params: "
                      f"{'whole toy model' if model == 'toy' else 'first four transformer blocks' if model.startswith('ViT') else 'attention + norms of the first decoder layer'}), {len(tasks)} tasks{upcast}, "
                      f"oracle/svd_hybrid_ref.py (torch-eager CPU restatement of the reference path; its full-feature "
                      f"k-means excluded), {dt:.2f} s per pass"}, dt, n_params


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, dt, n_params = cpu_baseline(args.workload, steps=args.steps, warmup=min(args.warmup, 1))
    cfg, tasks, model, p = _make_cfg(args.workload)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak" if (WORKLOAD_EXTRA.get(args.workload, {}).get("shard_world") or args.replicas) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": _workload_string(args.workload, getattr(args, "family", "parity")),
                       "placement": "host cores of the box; every step = one pass over the bounded sample named in "
                                    "cpu_baseline.sample"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(json.dumps(line))


def _bind_to_gpu_numa_node(gpu_index: int):
    """One process per GPU: run this rank's host threads on the CPUs next to its GPU, so that the pinned staging
    buffers of the end-to-end measurement are first-touched on the local NUMA node (8 ranks otherwise pull their
    13 GB per merge across the socket interconnect).  Best effort: silently skipped when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


def run_ours(args):
    import torch
    import torch.distributed as dist
    from svd_quantization_task_merging_b200 import _native, sharding, synth
    from svd_quantization_task_merging_b200.engine import MergeJob, pack_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _native.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg, tasks, model, p = _make_cfg(args.workload)
    shapes_all = synth.model_shapes(model)
    N = len(tasks)
    extra = WORKLOAD_EXTRA.get(args.workload, {})
    in_dtype = getattr(torch, extra.get("dtype", "float32"))
    cost = {k: int(np.prod(v)) * (N + 1) for k, v in shapes_all.items()}
    # How the N ranks split the work (north_star: "the model is partitioned per parameter across the GPUs"):
    #   strong  : ONE model, LPT-partitioned over the ranks; every rank holds, uploads and merges only its shard
    #   shard   : weak scaling of a capacity workload: rank r merges shard r of a fixed 8-way partition (Llama-3-8B:
    #             8 ranks = the whole 8.03 B-parameter model)
    #   replicas: weak scaling, every rank merges its own copy of the model (round-1 behaviour; --replicas)
    if extra.get("shard_world"):
        mode, sw = "shard", extra["shard_world"]
        owner = sharding.lpt_partition(cost, sw)
        mine = (args.shard_rank + rank) % sw
        shard_note = f"shard {mine} of the {sw}-way LPT partition per rank"
    elif args.replicas or world == 1:
        mode, owner, mine = ("replicas" if world > 1 else "single"), {k: 0 for k in shapes_all}, 0
        shard_note = "per-GPU copy of the workload" if world > 1 else "whole model on one GPU"
    else:
        mode, mine = "strong", rank
        owner = sharding.lpt_partition(cost, world)
        shard_note = f"ONE model LPT-partitioned by parameter over {world} ranks (strong scaling)"
    shapes = type(shapes_all)((k, v) for k, v in shapes_all.items() if owner[k] == mine)
    n_params = synth.total_params(shapes)
    # synthetic random-init checkpoints, resident in HBM before the timed region.  Every tensor has its own random
    # stream (per_tensor), so a rank that generates only its shard holds exactly the tensors of the whole model.
    seed = 1234 + (rank if mode == "replicas" else 0)
    gen = dict(family=args.family, seed=seed, device=str(dev), dtype=in_dtype, per_tensor=True)
    base, fts = synth.make_checkpoints(shapes, tasks, **gen)
    masks = synth.make_masks(shapes, tasks, p, seed=4321 + seed, device=str(dev), per_tensor=True) if p is not None else None
    perf = synth.performance_table(tasks) if cfg.svd_weighting == "performance" else None
    job = MergeJob(base, fts, masks, cfg, str(dev), performance=perf, diagnostics=False)
    collectives = []
    if mode == "strong" and job.cluster_mode:
        job.gram_reduce_hook = sharding.allreduce_gram          # N x N fp64 whole-model Gram, on the job's side stream
        collectives.append("all_reduce(8x8 fp64 whole-model Gram)")
    launches_per_step = job.gpu_launches
    torch.cuda.synchronize(dev)

    # per-parameter records (rank, k, energy, masked size) of all ranks: one flat all-gather per step (device side)
    rec_gather = None
    if mode == "strong":
        by_rank = [sum(1 for q in owner.values() if q == r) for r in range(world)]
        pmax = max(by_rank)
        rec_out = torch.empty(world * pmax, 16, dtype=torch.float64, device=dev)
        rec_buf = torch.zeros(pmax, 16, dtype=torch.float64, device=dev)
        collectives.append("all_gather_into_tensor(per-parameter records, fp64 [P,16])")

        def rec_gather():
            o = 0
            for g in job.groups.values():
                P = len(g.names)
                rec_buf[o: o + P, :8] = g.t["info"]
                rec_buf[o: o + P, 8] = g.t["dm"]
                rec_buf[o: o + P, 9:13] = g.t["scal"]
                o += P
            dist.all_gather_into_tensor(rec_out, rec_buf)

    def step(record_events=False):
        job.run(record_events=record_events)
        if rec_gather is not None:
            rec_gather()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    t_begin = time.perf_counter()
    k_times = {"k1": 0.0, "k2": 0.0, "k3": 0.0}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    jobs_events = []
    ev0.record()
    for _ in range(args.steps):
        step(record_events=True)
        jobs_events.append(job._events)
    ev1.record()
    barrier()
    t_end = time.perf_counter()
    total_ms = ev0.elapsed_time(ev1)
    for ev in jobs_events:
        k_times["k1"] += ev["start"].elapsed_time(ev["k1"])
        k_times["k2"] += ev["k1"].elapsed_time(ev["k2"])
        k_times["k3"] += ev["k2"].elapsed_time(ev["k3"])
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    # parameters merged per step by all ranks together
    if world > 1:
        c = torch.tensor([n_params], dtype=torch.int64, device=dev)
        gathered_n = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(gathered_n, c)
        params_per_rank = [int(x.item()) for x in gathered_n]
    else:
        params_per_rank = [n_params]
    n_total = sum(params_per_rank)
    value = n_total / (ms_per_step * 1e-3)

    # secondary figures: the same merge with the per-task reconstruction diagnostics fused into
    # pass 2 (the reference's default svd_eval_reconstruction=True), and with the bases (U_high / U_low fp16,
    # mean) materialised in the artifact layout
    secondary = None
    if world == 1 and not args.no_secondary:
        try:                  # auxiliary figures: a failure here must not lose the headline line
            def timed(fn, n):
                # median of per-call device times after two warm-up calls (the first calls pay cudaMalloc)
                fn()
                fn()
                torch.cuda.synchronize(dev)
                times = []
                for _ in range(n):
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    fn()
                    a1.record()
                    torch.cuda.synchronize(dev)
                    times.append(a0.elapsed_time(a1))
                return float(np.median(times))
            jd = MergeJob(base, fts, masks, cfg, str(dev), performance=perf, diagnostics=True)
            ms_diag = timed(jd.run, 5)
            del jd
            torch.cuda.empty_cache()

            def with_bases():
                job._bases_done = False
                job._materialize_bases()
            job.run()
            ms_art = timed(with_bases, 5)
            job._basis_store = {}
            torch.cuda.empty_cache()
            # the reference's default settings: diagnostics AND stored artifacts; pass 2 then writes the bases itself
            # (svdq_reconstruct_merge_basis, up to 8 tasks) instead of a third pass over the inputs
            ms_both = ms_art_fused = None
            if not cfg.svd_include_noise and N <= 8:
                for with_diag in (True, False):
                    jb = MergeJob(base, fts, masks, cfg, str(dev), performance=perf, diagnostics=with_diag, materialize_bases=True)
                    if jb._fused_basis_buffers() is not None:
                        if with_diag:
                            ms_both = timed(jb.run, 5)
                        else:
                            ms_art_fused = timed(jb.run, 5)
                    del jb
                    torch.cuda.empty_cache()
            secondary = {"with_fused_diagnostics": {"ms_per_step": ms_diag, "value": n_params / (ms_diag * 1e-3)},
                         "basis_materialisation_extra_ms": ms_art,
                         "with_artifacts": {"ms_per_step": ms_per_step + ms_art,
                                            "value": n_params / ((ms_per_step + ms_art) * 1e-3),
                                            "note": "separate third pass (svdq_write_basis)"},
                         "with_artifacts_fused": None if ms_art_fused is None else
                         {"ms_per_step": ms_art_fused, "value": n_params / (ms_art_fused * 1e-3),
                          "note": "bases written by pass 2 itself (svdq_reconstruct_merge_basis without diagnostics)"},
                         "with_diagnostics_and_artifacts": None if ms_both is None else
                         {"ms_per_step": ms_both, "value": n_params / (ms_both * 1e-3),
                          "note": "bases written by pass 2 itself (reference defaults: svd_eval_reconstruction + svd_store_artifacts)"}}
        except Exception as exc:      # noqa: BLE001
            secondary = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()

    fetched = job._fetch()
    solved = sum(int((f["info"][:, 0] == 0).sum()) for f in fetched.values())
    if world > 1:
        s = torch.tensor([solved], dtype=torch.int64, device=dev)
        gathered = [torch.zeros_like(s) for _ in range(world)]
        dist.all_gather(gathered, s)
        solved_all = [int(x.item()) for x in gathered]
    else:
        solved_all = [solved]

    # strong scaling extras (outside the timed region): the optional all-gather of the merged shards over NVLink
    # (one padded all_gather_into_tensor), and a bit-identity check of the sharded result against ONE GPU merging
    # the whole model (rank 0 regenerates the full model; per-tensor streams make the shards identical)
    sharded = None
    if mode == "strong":
        sizes = [0] * world
        flat = next(iter(job.groups.values())).t["out"] if len(job.groups) == 1 else \
            torch.cat([g.t["out"] for g in job.groups.values()])
        sz = torch.tensor([flat.numel()], dtype=torch.int64, device=dev)
        gsz = [torch.zeros_like(sz) for _ in range(world)]
        dist.all_gather(gsz, sz)
        sizes = [int(x.item()) for x in gsz]
        allm = sharding.gather_merged(flat, sizes)          # warm-up
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            allm = sharding.gather_merged(flat, sizes)
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        equal, assign_equal, checked = None, None, 0
        assign = [job.cluster_assignments]
        if rank == 0:
            b_all, f_all = synth.make_checkpoints(shapes_all, tasks, **gen)
            m_all = synth.make_masks(shapes_all, tasks, p, seed=4321 + seed, device=str(dev), per_tensor=True) \
                if p is not None else None
            one = MergeJob(b_all, f_all, m_all, cfg, str(dev), performance=perf, diagnostics=False).run()
            ref = one.merged_state_dict()
            equal = True
            for r in range(world):
                names_r = sorted(k for k in shapes_all if owner[k] == r)
                off = 0
                for k in names_r:          # arena layout of rank r: sorted names, 64-element aligned
                    n = int(np.prod(shapes_all[k]))
                    got = allm[r, off: off + n].view(shapes_all[k])
                    equal = equal and bool(torch.equal(got, ref[k]))
                    checked += 1
                    off += (n + 63) // 64 * 64
            assign_equal = one.cluster_assignments == job.cluster_assignments
            del b_all, f_all, m_all, one, ref
        del allm
        torch.cuda.empty_cache()
        barrier()
        sharded = {"gather_merged_ms": float(tg.item()), "gather_merged_bytes_per_rank": int(max(sizes)) * 4 * world,
                   "equals_single_gpu_bitwise": equal, "tensors_compared": checked,
                   "cluster_assignments_equal": assign_equal}

    # roofline of the dominant kernel (algorithmic bytes per element x elements per launch / measured time);
    # this rank's launch processes this rank's elements
    has_masks = masks is not None
    es = torch.empty(0, dtype=in_dtype).element_size()
    bytes_k1 = n_params * ((N + 1) * es + (N + 1 / 8 if has_masks else 0))
    bytes_k3 = n_params * ((N + 1) * es + (1 / 8 if has_masks else 0) + 4)
    peak, peak_src = _measured_peak()
    k1_ms, k3_ms, k2_ms = (k_times[k] / args.steps for k in ("k1", "k3", "k2"))
    dom = "k1_tv_mask_gram" if k1_ms >= k3_ms else "k3_reconstruct_merge"
    dom_bytes, dom_ms = (bytes_k1, k1_ms) if k1_ms >= k3_ms else (bytes_k3, k3_ms)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            traffic = (json.load(open(tpath)).get(args.workload) or {}).get(dom)      # null for workloads without a capture
        except Exception:
            traffic = None
    whole_bytes = (bytes_k1 + bytes_k3) / n_params * n_total        # all ranks' algorithmic bytes per step
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": dom_ms,
                "kernels_ms_per_step": {"k1_tv_mask_gram": k1_ms, "k2_reduce+cluster+solve": k2_ms,
                                        "k3_reconstruct_merge": k3_ms},
                "whole_path": {"algorithmic_bytes_per_step": whole_bytes,
                               "achieved": whole_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": whole_bytes / (ms_per_step * 1e-3) / 1e9 / (peak * world)}}

    # end to end through the public API with HOST buffers: H2D of every input + D2H of the merged model per step
    e2e = None
    if not args.no_e2e:
        from svd_quantization_task_merging_b200.engine import merge_state_dicts
        h_base = pack_state_dict({k: v.cpu() for k, v in base.items()}, pin=True)
        h_fts = {t: pack_state_dict({k: v.cpu() for k, v in fts[t].items()}, pin=True) for t in tasks}
        h_masks = ({t: pack_state_dict({k: v.cpu() for k, v in masks[t].items()}, pin=True) for t in tasks}
                   if masks is not None else None)
        del job, base, fts, masks
        torch.cuda.empty_cache()
        h2d = d2h = 0
        e_steps = max(1, min(args.steps, 3))

        def e2e_call():
            if mode == "strong":      # public sharded API: per-rank shard upload, Gram all-reduce, flat record gather
                return sharding.merge_state_dicts_sharded(h_base, h_fts, h_masks, cfg, str(dev), owner=owner,
                                                          performance=perf, diagnostics=False, to_host="reuse")
            return merge_state_dicts(h_base, h_fts, h_masks, cfg, str(dev), performance=perf, diagnostics=False,
                                     to_host="reuse")
        e2e_call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            r = e2e_call()
            h2d = r["job"].h2d_bytes
            d2h = sum(g.t["out"].numel() * 4 for g in r["job"].groups.values())
            del r
        barrier()
        dt = (time.perf_counter() - t0) / e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        hb = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(hb)
        # one more call with per-phase synchronisation: where the end-to-end time goes on rank 0 (not part of the timing)
        os.environ["SVDQ_PROFILE"] = "1"
        t0 = time.perf_counter()
        r = e2e_call()
        phases = dict(r["job"].timing)
        phases["total_s"] = time.perf_counter() - t0
        del r
        os.environ["SVDQ_PROFILE"] = "0"
        barrier()
        e2e = {"value": n_total / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(hb[0].item()),
               "d2h_bytes_per_step": int(hb[1].item()), "ms_per_step": float(tt.item()) * 1e3, "steps": e_steps,
               "bytes_are": "summed over all ranks", "phases_rank0_synchronised": phases}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                cpu, _, _ = cpu_baseline(args.workload, steps=1, warmup=0)
            except Exception as exc:      # noqa: BLE001 -- reported baseline only: keep the measured line
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "",
                       "error": f"{type(exc).__name__}: {exc}"[:300]}
        dt_name = {"float32": "f32", "bfloat16": "bf16", "float16": "f16"}[str(in_dtype).split(".")[-1]]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak" if mode in ("shard", "replicas") else "strong",
                "vs_baseline": None, "dtype": dt_name, "data": "synthetic",
                "config": {"workload": _workload_string(args.workload, args.family),
                           "placement": shard_note,
                           "l2": f"inputs ({bytes_k1 / 1e9:.1f} GB per rank and step) "
                                 + ("are larger than the 126 MB L2; no flush needed" if bytes_k1 > 4e8 else
                                    "-- see DESIGN.md section 7 on L2 residency at high rank counts"),
                           "diagnostics_fused": False, "artifacts_materialised": False,
                           "cluster_backend": "kmeans (reference procedure, svdq_host_kmeans)",
                           "parallelism": f"{mode}: {world} GPU(s), parameter-sharded, no data-path collective; "
                                          f"collectives per step: {collectives if collectives else 'none'}"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
                "gpu_launches": launches_per_step * args.steps, "params_per_rank": params_per_rank,
                "params_with_basis_per_rank": solved_all, "sharded": sharded, "secondary": secondary}
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """The bench prints ONE JSON line on stdout.  Native libraries write to file descriptor 1 directly (NCCL's
    version banner), so fd 1 is pointed at stderr for the duration of the run and the line goes to the saved fd."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(text: str):
    if _REAL_STDOUT is None:
        print(text)
        return
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vit-l-14-cluster", choices=sorted(WORKLOADS))
    ap.add_argument("--shard-rank", type=int, default=0, help="which shard rank 0 takes (sharded workloads)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--secondary", action="store_true", help="(default at 1 GPU) time the fused-diagnostics and artifact variants")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--replicas", action="store_true",
                    help="N > 1: weak scaling with one model copy per GPU instead of one model sharded over the GPUs")
    ap.add_argument("--family", default="parity", choices=["parity", "throughput"],
                    help="synthetic input family (SURVEY 8d): decaying spectrum | iid")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
