"""Weighted averaging in coefficient space and reconstruction.  Mirror of src/svd_hybrid/merge.py:61-626.

These operator-level functions work on the reference's artifact structures (bases + compressed
coefficients) and serve the artifact reload path; the pipeline itself merges with the fused K3
kernel (engine.MergeJob), which folds averaging weights, reconstruction, mask scatter and
``base + delta`` into one pass."""
from typing import Dict, List, Optional, Tuple

import torch

from .. import _native
from . import _ops
from .rtvq import RTVQQuantizer


def dequantize_and_average(compressed_coeffs: Dict[str, Dict], weights: Dict[str, float], quantizer: RTVQQuantizer,
                           region: str = "masked", device: str = "cpu") -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """merge.py:61-141: sorted task order; weights renormalised over the tasks that have the region."""
    names = sorted(compressed_coeffs.keys())
    hi, lo, w = [], [], []
    for n in names:
        art = compressed_coeffs[n]
        if art is None or art.get(region) is None:
            continue
        reg = art[region]
        hi.append(reg["c_high_fp16"].to(device).float())
        lo.append(quantizer.dequantize(reg["c_low_quant"], device=device).float())
        w.append(weights.get(n, 1.0 / len(names)))
    if not hi:
        return None, None
    tot = sum(w)
    wt = torch.tensor([x / tot for x in w], device=device, dtype=torch.float32).view(-1, 1)
    return (torch.stack(hi, 0) * wt).sum(0), (torch.stack(lo, 0) * wt).sum(0)


def reconstruct_from_coefficients(avg_c_high: torch.Tensor, avg_c_low: torch.Tensor, U_high: torch.Tensor,
                                  U_low: torch.Tensor, device: str = "cpu", mean: Optional[torch.Tensor] = None) -> torch.Tensor:
    """U_high c_high + U_low c_low (+ mean) in fp32 (merge.py:144-194) on the device (K14, svdq_basis_expand)."""
    return _ops.expand(avg_c_high, avg_c_low, U_high, U_low, mean=mean).to(device)


def merge_parameter(param_name: str, compressed_params: Dict[str, Dict], basis: Dict, weights: Dict[str, float],
                    quantizer: RTVQQuantizer, original_shape: torch.Size, mask: Optional[torch.Tensor] = None,
                    include_noise: bool = False, noise_shrink: float = 1.0, device: str = "cpu") -> torch.Tensor:
    """merge.py:197-301."""
    from .mask_loader import reconstruct_from_masked

    def region(reg_basis, reg_name):
        if reg_basis is None:
            return None
        c_hi, c_lo = dequantize_and_average(compressed_params, weights, quantizer, region=reg_name, device=device)
        if c_hi is None:
            return None
        return reconstruct_from_coefficients(c_hi, c_lo, reg_basis["U_high"], reg_basis["U_low"], device=device,
                                             mean=reg_basis.get("mean"))

    merged_masked = region(basis.get("masked"), "masked")
    merged_unmasked = None
    if include_noise:
        merged_unmasked = region(basis.get("noise"), "unmasked")
        if merged_unmasked is not None:
            merged_unmasked = merged_unmasked * noise_shrink
    if mask is not None and merged_masked is not None:
        return reconstruct_from_masked(merged_masked, merged_unmasked, mask, original_shape)
    if merged_masked is not None:
        return merged_masked.view(original_shape)
    return torch.zeros(original_shape, device=device)


def merge_all_parameters(compressed_all: Dict[str, Dict[str, Dict]], bases: Dict[str, Dict], masks: Dict[str, torch.Tensor],
                         weights: Dict[str, float], original_shapes: Dict[str, torch.Size], config, device: str = "cpu",
                         verbose: bool = True) -> Dict[str, torch.Tensor]:
    """merge.py:304-426: every parameter of ``compressed_all`` in sorted order -- as ONE batched launch over the stored
    bases (K11, svdq_reload_merge: pass 2 with U read instead of rebuilt) instead of a matvec pair per parameter.
    The coefficient averages (a few numbers per parameter, merge.py:61-141) are formed on the host."""
    import numpy as np
    _native.require_cuda()
    g = torch.device("cuda")
    quantizer = RTVQQuantizer(num_bits=config.svd_low_bits, num_stages=config.svd_rtvq_stages)
    names = sorted(compressed_all.keys())
    P = len(names)
    if P == 0:
        return {}
    te = 16384
    include_noise = bool(getattr(config, "svd_include_noise", False))
    shrink = float(getattr(config, "svd_noise_shrink", 1.0))
    numel = np.asarray([int(torch.Size(original_shapes[n]).numel()) for n in names], np.int64)

    def region_tables(basis_key, coeff_region):
        """-> (kr [P][2], U_high / U_low / mean tensors on the GPU per parameter, cbar rows)"""
        kr = np.zeros((P, 2), np.int32)
        uh, ul, mn, cb = [None] * P, [None] * P, [None] * P, [None] * P
        for p, n in enumerate(names):
            b = (bases.get(n) or {}).get(basis_key)
            if b is None:
                continue
            c_hi, c_lo = dequantize_and_average(compressed_all[n], weights, quantizer, region=coeff_region, device="cpu")
            if c_hi is None:
                continue
            kr[p] = (c_hi.numel(), c_hi.numel() + c_lo.numel())
            uh[p], ul[p] = b["U_high"], b["U_low"]
            mn[p] = b.get("mean")
            cb[p] = torch.cat([c_hi.float().cpu(), c_lo.float().cpu()])
        return kr, uh, ul, mn, cb

    regions = [region_tables("masked", "masked")]
    if include_noise:
        regions.append(region_tables("noise", "unmasked"))
    nt = max(1, max(int(r[0][:, 1].max()) for r in regions))
    if nt > 32:
        raise ValueError("more than 32 basis directions per parameter")
    # packed combined masks (bits) -> device; parameters without a mask keep every row
    has_mask = np.asarray([1 if masks.get(n) is not None else 0 for n in names], np.uint8)
    pm_off = np.zeros(P, np.int64)
    words = 0
    for p in range(P):
        pm_off[p] = words
        if has_mask[p]:
            words += ((int(numel[p]) + 31) // 32 + 3) // 4 * 4
    stage = torch.zeros(max(words, 4) * 4, dtype=torch.uint8).pin_memory()
    src, cnt, dst, keep = [], [], [], []
    on_dev = {}
    for p, n in enumerate(names):
        if not has_mask[p]:
            continue
        m = masks[n]
        if tuple(m.shape) != tuple(torch.Size(original_shapes[n])):
            raise ValueError(f"mask shape {tuple(m.shape)} does not match parameter shape {tuple(original_shapes[n])}")
        m = m.detach().to("cpu", torch.bool).contiguous()
        keep.append(m)
        src.append(m.data_ptr()); cnt.append(m.numel()); dst.append(stage.data_ptr() + 4 * int(pm_off[p]))
    if src:
        a_src, a_cnt, a_dst = (np.asarray(x, np.int64) for x in (src, cnt, dst))
        _native.call("svdq_host_pack_mask_batch", a_src.ctypes.data, a_cnt.ctypes.data, a_dst.ctypes.data, len(src), 8)
    packed = stage.to(g, non_blocking=True).view(torch.int32)
    # tile tables
    tiles_per = (numel + te - 1) // te
    tile_begin = np.zeros(P + 1, np.int64)
    np.cumsum(tiles_per, out=tile_begin[1:])
    n_tiles = int(tile_begin[-1])
    tile_param = np.repeat(np.arange(P, dtype=np.int32), tiles_per)
    tile_local = (np.arange(n_tiles, dtype=np.int64) - np.repeat(tile_begin[:-1], tiles_per)).astype(np.int32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(g)      # noqa: E731
    out_off = np.zeros(P + 1, np.int64)
    np.cumsum((numel + 63) // 64 * 64, out=out_off[1:])
    out = torch.empty(max(int(out_off[-1]), 1), dtype=torch.float32, device=g)
    t = dict(numel=dev(numel), tile_param=dev(tile_param), tile_local=dev(tile_local), tile_begin=dev(tile_begin),
             pm_off=dev(pm_off), has_mask=dev(has_mask), count=torch.zeros(max(n_tiles, 1), dtype=torch.int32, device=g),
             row_off=torch.zeros(max(n_tiles, 1), dtype=torch.int64, device=g),
             optr=dev(np.asarray([out.data_ptr() + 4 * int(o) for o in out_off[:-1]], np.int64)))
    st = _native.stream_ptr()
    ptr = lambda x: None if x is None else x.data_ptr()      # noqa: E731
    with torch.cuda.device(g):
        _native.call("svdq_mask_tile_counts", n_tiles, te, ptr(t["numel"]), ptr(t["tile_param"]), ptr(t["tile_local"]),
                     ptr(t["pm_off"]), ptr(t["has_mask"]), ptr(packed), ptr(t["count"]), st)
        for reg, (kr, uh, ul, mn, cb) in enumerate(regions):
            fp16 = all(u.dtype == torch.float16 for u in uh if u is not None)
            udt = torch.float16 if fp16 else torch.float32
            hold = []

            def table(xs, dt):
                ptrs = np.zeros(P, np.int64)
                for p, x in enumerate(xs):
                    if x is None or x.numel() == 0:
                        continue
                    y = x.detach().to(g, dt).contiguous()
                    hold.append(y)
                    ptrs[p] = y.data_ptr()
                return dev(ptrs)
            uh_t, ul_t = table(uh, udt), table(ul, udt)
            mn_t = table(mn, torch.float32) if any(m is not None for m in mn) else None
            cbar = np.zeros((P, nt), np.float32)
            for p, c in enumerate(cb):
                if c is not None:
                    cbar[p, : c.numel()] = c.numpy()
            kr_d, cbar_d = dev(kr), dev(cbar)               # named: they must outlive the launch
            _native.call("svdq_basis_offsets", P, reg, te, ptr(t["count"]), ptr(t["tile_begin"]), ptr(t["numel"]),
                         ptr(t["row_off"]), st)
            _native.call("svdq_reload_merge", int(fp16), nt, reg, shrink if reg == 1 else 1.0, n_tiles, te,
                         ptr(t["numel"]), ptr(t["tile_param"]), ptr(t["tile_local"]), ptr(t["pm_off"]), ptr(t["has_mask"]),
                         ptr(packed), ptr(kr_d), ptr(uh_t), ptr(ul_t), ptr(mn_t), ptr(cbar_d), ptr(t["row_off"]),
                         ptr(t["optr"]), st)
            torch.cuda.current_stream().synchronize()      # `hold` (the staged bases) may go once the launch has run
    res = {}
    for p, n in enumerate(names):
        res[n] = out[int(out_off[p]): int(out_off[p]) + int(numel[p])].view(torch.Size(original_shapes[n])).to(device)
    if verbose:
        total = sum(d.norm().item() ** 2 for d in res.values()) ** 0.5
        print(f"   merged {len(res)} parameters, total delta norm {total:.6f}")
    return res


def apply_merged_deltas(base_state_dict: Dict[str, torch.Tensor], merged_deltas: Dict[str, torch.Tensor],
                        device: str = "cpu", verbose: bool = True) -> Dict[str, torch.Tensor]:
    """merged[p] = base[p] + delta[p]; other keys are cloned (merge.py:429-552)."""
    out = {}
    for name, b in base_state_dict.items():
        out[name] = b + merged_deltas[name].to(b.device) if name in merged_deltas else b.clone()
    if verbose:
        print(f"   applied {sum(1 for n in base_state_dict if n in merged_deltas)} deltas to {len(out)} parameters")
    return out


def merge_with_clustering(compressed_all: Dict[str, Dict[str, Dict]], bases: Dict[str, Dict], masks: Dict[str, torch.Tensor],
                          weights: Dict[str, float], cluster_assignments: Dict[str, int],
                          original_shapes: Dict[str, torch.Size], config, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """merge.py:555-626: merge inside each cluster, then softmax(mean member weight)-average the clusters."""
    from .clustering import get_cluster_members, merge_cluster_results
    clusters = get_cluster_members(cluster_assignments)
    per_cluster, score = {}, {}
    for cid, members in clusters.items():
        cw = {n: weights.get(n, 1.0) for n in members}
        tot = sum(cw.values())
        cw = {n: v / tot for n, v in cw.items()}
        sub = {p: {n: art[n] for n in members if n in art} for p, art in compressed_all.items()}
        per_cluster[cid] = merge_all_parameters(sub, bases, masks, cw, original_shapes, config, device, verbose=False)
        score[cid] = sum(weights.get(n, 1.0) for n in members) / len(members)
    return merge_cluster_results(per_cluster, score, device)
