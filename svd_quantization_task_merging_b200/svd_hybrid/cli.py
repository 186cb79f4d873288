"""Orchestrator and command line of the SVD-Hybrid merge.  Mirror of src/svd_hybrid/cli.py:73-961.

``run_svd_hybrid_pipeline(config)`` keeps the reference's contract (same inputs on disk, same result
dict, same files written) but executes steps 1-9 as the fused GPU path (engine.MergeJob): checkpoints
and masks are read with torch.load, staged to the device, merged by the K1/K2/K3 kernels, and the
artifacts are written in the reference layout."""
import argparse
import json
import os
from typing import Dict

import torch

from .config import SVDHybridConfig
from .diagnostics import compression_statistics_from_sizes, print_detailed_compression_report, print_diagnostics_summary
from .mask_loader import load_task_masks
from .storage import save_all_artifacts, save_combined_masks, save_merged_model
from .task_vector_loader import get_task_checkpoint_paths, load_checkpoint


def run_svd_hybrid_pipeline(config: SVDHybridConfig, verbose: bool = True) -> Dict:
    """cli.py:73-778 -> {"merged_state_dict", "diagnostics", "bases", "compressed"}."""
    from .. import _native
    from ..engine import MergeJob
    _native.require_cuda()                       # no CPU fallback: without libsvdq.so or a CUDA device this raises
    # The reference's `device` says where torch computes AND where the results live (its own integration tests pass
    # device="cpu", tests/test_integration.py:77,165).  Here the arithmetic always runs on the GPU; device="cpu" keeps
    # its second meaning -- the returned state dict lives in host memory -- and is said out loud, not honoured silently.
    want_host = str(config.device).startswith("cpu")
    device = config.device if str(config.device).startswith("cuda") else "cuda"

    def say(msg):
        if verbose:
            print(msg)

    if want_host:
        print('[svd-hybrid] device="cpu" requested: the SVD-Hybrid path runs on the GPU in this build (no CPU '
              "implementation); results are returned in host memory")

    say(f"[svd-hybrid] base model: {config.base_model_path}")
    base = load_checkpoint(config.base_model_path, device="cpu")                       # step 0 (cli.py:146)
    paths = get_task_checkpoint_paths(config.checkpoint_dir, config.tasks)             # step 1 (cli.py:164)
    finetuned = {t: load_checkpoint(p, device="cpu") for t, p in paths.items()}
    task_masks = None
    if config.mask_dir and os.path.exists(config.mask_dir):                            # step 2 (cli.py:201-209)
        task_masks = load_task_masks(config.mask_dir, config.tasks, device="cpu", reference_state_dict=base,
                                     verbose=verbose)
    say(f"[svd-hybrid] {len(config.tasks)} tasks, {len(base)} base tensors, masks: {'yes' if task_masks else 'no'}")

    # checkpoint ingest (steps 0-1 end here): pageable torch.load dicts -> pinned staging -> device, the packing of
    # state dict t+1 overlapped with the DMA of state dict t (engine.upload_state_dicts)
    from ..engine import upload_state_dicts
    order = [t for t in config.tasks if t in finetuned]
    staged = upload_state_dicts([base] + [finetuned[t] for t in order], device)
    base_d, finetuned_d = staged[0], dict(zip(order, staged[1:]))
    job = MergeJob(base_d, finetuned_d, task_masks, config, device, materialize_bases=bool(config.svd_store_artifacts))
    job.run()                                                                          # steps 3-9 on the GPU
    res = job.results(to_host=want_host)
    merged, bases, compressed, diagnostics = (res["merged_state_dict"], res["bases"], res["compressed"],
                                              res["diagnostics"])
    weights = job.weights
    say("[svd-hybrid] task weights: " + ", ".join(f"{t}={w:.4f}" for t, w in sorted(weights.items())))

    # step 5b: compression report (cli.py:447-453) from sizes -- no need to materialise the bases for it
    numel = {n: int(torch.Size(s).numel()) for n, s in job.shapes.items()}
    sizes = {}
    for name in bases:
        m = bases.meta(name)
        dt, p = bases._index[name]
        present = int(job.groups[dt].host["present"][p])
        sizes[name] = {"k": m["k"], "r": m["r"], "D": m["D"], "n_tasks": bin(present).count("1")}
    n_with = {n: bin(int(job.groups[dt].host["present"][p])).count("1")
              for dt, g in job.groups.items() for p, n in enumerate(g.names)}
    stats = compression_statistics_from_sizes(numel, n_with, sizes, len(config.tasks), config.svd_low_bits,
                                              config.svd_rtvq_stages, list(config.tasks))
    if verbose:
        print_detailed_compression_report(stats, config)
        if config.svd_eval_reconstruction:
            print_diagnostics_summary(diagnostics)

    if config.svd_store_artifacts:                                                     # cli.py:727-735
        save_all_artifacts(bases, compressed, diagnostics, config, config.artifact_dir)
        # additive to the reference layout: the combined tall masks, without which a masked run cannot be re-merged
        # from its artifacts (the reference's reload passes no masks, reload.py:204, and fails on the scatter)
        cm = job.combined_masks()
        if cm:
            save_combined_masks(cm, config.artifact_dir)
    save_merged_model(merged, config.output_dir)                                       # cli.py:740
    os.makedirs(config.output_dir, exist_ok=True)
    with open(os.path.join(config.output_dir, "weights.json"), "w") as f:
        json.dump(weights, f, indent=2)
    if job.cluster_assignments:
        with open(os.path.join(config.output_dir, "clusters.json"), "w") as f:
            json.dump(job.cluster_assignments, f, indent=2)
    say(f"[svd-hybrid] merged model written to {config.output_dir}")
    return {"merged_state_dict": merged, "diagnostics": diagnostics, "bases": bases, "compressed": compressed}


def parse_args(argv=None):
    """Same flags and defaults as cli.py:781-875."""
    p = argparse.ArgumentParser(description="SVD-Hybrid merging method combining Tall Masks and TVQ")
    p.add_argument("--config", type=str, default=None, help="Path to JSON config file (overrides command-line args)")
    p.add_argument("--quantize-config", type=str, default=None, help="Path to quantization config JSON")
    p.add_argument("--load-config", type=str, default=None, help="Path to loading config JSON")
    p.add_argument("--tasks", nargs="+", help="List of task identifiers")
    p.add_argument("--model", type=str, default="ViT-B-32", help="Model identifier (e.g., ViT-B-32)")
    p.add_argument("--checkpoint-dir", type=str, help="Directory containing task checkpoints")
    p.add_argument("--base-model-path", type=str, help="Path to base model checkpoint")
    p.add_argument("--mask-dir", type=str, default="", help="Directory containing tall masks")
    p.add_argument("--load-tv-type", type=str, default=None,
                   choices=["standard", "quantized", "quantized_finetuned", "quantized_base_and_tv"],
                   help="Type of task vector to load")
    p.add_argument("--load-task-bits", type=int, default=8, help="Bits for task vector quantization when loading")
    p.add_argument("--load-base-bits", type=int, default=8, help="Bits for base model quantization when loading")
    p.add_argument("--energy-threshold", type=float, default=0.95, help="Energy retention threshold for rank selection")
    p.add_argument("--max-rank", type=int, default=64, help="Maximum rank cap")
    p.add_argument("--center", action="store_true", default=True, help="Center task matrix before SVD")
    p.add_argument("--no-center", action="store_false", dest="center", help="Don't center task matrix")
    p.add_argument("--fp16", action="store_true", default=True, help="Use FP16 for bases")
    p.add_argument("--no-fp16", action="store_false", dest="fp16", help="Use FP32 for bases")
    p.add_argument("--low-bits", type=int, default=4, help="Bits for low-energy coefficient quantization")
    p.add_argument("--rtvq-stages", type=int, default=2, help="Number of RTVQ refinement stages")
    p.add_argument("--mask-strategy", type=str, default="union", choices=["union", "intersection", "majority"],
                   help="Mask combination strategy")
    p.add_argument("--include-noise", action="store_true", help="Process unmasked (noise) region")
    p.add_argument("--noise-shrink", type=float, default=0.5, help="Shrinkage factor for noise region")
    p.add_argument("--weighting", type=str, default="uniform", choices=["uniform", "performance", "cluster"],
                   help="Task weighting strategy")
    p.add_argument("--performance-file", type=str, default=None, help="Path to performance metrics JSON file")
    p.add_argument("--temperature", type=float, default=5.0, help="Temperature for performance-based weighting")
    p.add_argument("--cluster-k", type=int, default=2, help="Number of clusters for cluster-based weighting")
    p.add_argument("--store-artifacts", action="store_true", help="Store compression artifacts")
    p.add_argument("--eval-reconstruction", action="store_true", default=True, help="Evaluate reconstruction error")
    p.add_argument("--no-eval-reconstruction", action="store_false", dest="eval_reconstruction",
                   help="Skip reconstruction evaluation")
    p.add_argument("--output-dir", type=str, default="./svd_hybrid_output", help="Output directory for merged model")
    p.add_argument("--artifact-dir", type=str, default="./artifacts", help="Directory for artifact storage")
    p.add_argument("--device", type=str, default="cuda", help="Device to use (cuda)")
    return p.parse_args(argv)


def config_from_args(args) -> SVDHybridConfig:
    """JSON overlays as cli.py:883-952: only tasks / checkpoint_dir / base_model_path are honoured."""
    over = {}
    if args.config:
        with open(args.config) as f:
            over = json.load(f)
    for path, section in ((args.quantize_config, "quantization"), (args.load_config, "loading")):
        if not path:
            continue
        with open(path) as f:
            extra = json.load(f)
        if section in extra:
            over.update(extra[section])
        if "tasks" in extra and not args.tasks:
            over["tasks"] = extra["tasks"]
        if "checkpoints" in extra:
            over.update(extra["checkpoints"])
    tasks = over.get("tasks", args.tasks)
    ckpt = over.get("checkpoint_dir", args.checkpoint_dir)
    base = over.get("base_model_path", args.base_model_path)
    if not tasks:
        raise ValueError("--tasks must be specified either via command-line or config file")
    if not ckpt:
        raise ValueError("--checkpoint-dir must be specified either via command-line or config file")
    if not base:
        raise ValueError("--base-model-path must be specified either via command-line or config file")
    return SVDHybridConfig(
        tasks=tasks, model=args.model, checkpoint_dir=ckpt, base_model_path=base, mask_dir=args.mask_dir,
        svd_energy_threshold=args.energy_threshold, svd_max_rank=args.max_rank, svd_center=args.center,
        svd_fp16=args.fp16, svd_low_bits=args.low_bits, svd_rtvq_stages=args.rtvq_stages,
        svd_mask_strategy=args.mask_strategy, svd_include_noise=args.include_noise, svd_noise_shrink=args.noise_shrink,
        svd_weighting=args.weighting, performance_file=args.performance_file,
        svd_weighting_temperature=args.temperature, svd_cluster_k=args.cluster_k,
        svd_store_artifacts=args.store_artifacts, svd_eval_reconstruction=args.eval_reconstruction,
        output_dir=args.output_dir, artifact_dir=args.artifact_dir, device=args.device)


def main(argv=None):
    return run_svd_hybrid_pipeline(config_from_args(parse_args(argv)))


if __name__ == "__main__":
    main()
