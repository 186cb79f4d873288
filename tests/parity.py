"""Shared helpers of the parity tests: run the oracle and the CUDA path on the same seeded inputs
and compare them the way SURVEY.md section 8c prescribes (sign alignment, subspace angles,
NaN-aware equality).  The oracle is only ever the checker here."""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

from oracle import svd_hybrid_ref as R
from svd_quantization_task_merging_b200 import synth
from svd_quantization_task_merging_b200.svd_hybrid.config import SVDHybridConfig

# tolerances (SURVEY.md 8c "stated tolerances to adopt")
TOL_S = 2e-6            # singular values, relative to sigma_1
TOL_S_GRAM = 2e-9       # ... for sigma_j << sigma_1: |d sigma_j| <= TOL_S_GRAM * sigma_1^2 / sigma_j (Gram route)
TOL_MERGED = 1e-5       # merged weights, relative L2 per parameter (sign-aligned), fp32 bases, when the stored
                        # artifacts (fp16 c_high bits, RTVQ codes) are identical to the oracle's
TOL_MERGED_FP16B = 1e-4 # ... with fp16 bases: LAPACK's fp32 left vectors carry an error of ~eps * sigma_1 / sigma_j
                        # (2e-6 relative for the weakest direction of the test spectra), so 1-2 % of the basis
                        # entries land on the other side of an fp16 rounding boundary (one fp16 ulp = 5e-4
                        # relative on that entry) => sqrt(0.015) * 4e-4 ~ 5e-5 relative L2 on small parameters
TOL_MERGED_FLIP = 2e-3  # ... when a coefficient sat on a rounding boundary and one fp16 value / code
                        # differs by one step: bounded by the fp16 / RTVQ step itself
TOL_COEF = 1e-5         # raw coefficients, abs relative to ||c||_inf
TOL_DIAG = 1e-4         # diagnostics floats, relative
TOL_ANGLE = 1e-5        # sine of the largest principal angle between spans

# observed equality rates of every compare_run call of the session (written to gpurun_out/parity_rates.json by
# tests/conftest.py; the committed copy is profiles/parity_rates.json)
RATES = []

MEDIUM_SHAPES = OrderedDict([
    ("blk.attn.weight", (300, 70)), ("blk.bias", (4099,)), ("conv.weight", (17, 33, 5)),
    ("wide.weight", (1, 1, 40000)), ("ln.weight", (768,)), ("tiny", (7,)), ("scalar_like", (1,)),
])


def make_cfgs(tasks, **kw):
    """-> (RefConfig for the oracle, SVDHybridConfig for the CUDA path) with identical settings."""
    perf = kw.pop("performance", None)
    ref = R.RefConfig(tasks=list(tasks), performance=perf, **kw)
    cfg = SVDHybridConfig(tasks=list(tasks), svd_store_artifacts=False, **kw)
    return ref, cfg


def run_both(shapes, n_tasks: int, mask_p: Optional[float] = None, family: str = "parity", seed: int = 1234,
             sign_align: bool = True, dtype=torch.float32, device="cuda", **cfg_kw):
    from svd_quantization_task_merging_b200.engine import merge_state_dicts
    tasks = synth.task_names(n_tasks)
    base, fts = synth.make_checkpoints(shapes, tasks, family=family, seed=seed, dtype=dtype)
    masks = synth.make_masks(shapes, tasks, mask_p, seed=seed + 1) if mask_p is not None else None
    perf = synth.performance_table(tasks) if cfg_kw.get("svd_weighting") == "performance" else None
    ref_cfg, cfg = make_cfgs(tasks, performance=perf, **cfg_kw)
    ref = R.run_reference_path(base, fts, masks, ref_cfg)
    ref["_base"] = base
    sign_ref = {p: b["Vh"] for p, b in ref["bases"].items()} if sign_align else None
    sign_ref_noise = {p: b["Vh"] for p, b in ref["bases_noise"].items()} if sign_align else None
    res = merge_state_dicts(base, fts, masks, cfg, device, sign_ref=sign_ref, sign_ref_noise=sign_ref_noise,
                            performance=perf,
                            cluster_assignments=ref["cluster_assignments"] if cfg.svd_weighting == "cluster" else None)
    return ref, res, (base, fts, masks, tasks)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / den if den > 0 else (a - b).norm().item()


def nan_positions_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    return bool(torch.equal(torch.isnan(a.cpu()), torch.isnan(b.cpu())))


def max_principal_sine(A: torch.Tensor, B: torch.Tensor) -> float:
    """sine of the largest principal angle between span(A) and span(B) (columns)."""
    A, B = A.double().cpu(), B.double().cpu()
    if A.shape[1] == 0 and B.shape[1] == 0:
        return 0.0
    qa, _ = torch.linalg.qr(A)
    qb, _ = torch.linalg.qr(B)
    resid = qb - qa @ (qa.T @ qb)
    return float(torch.linalg.matrix_norm(resid, ord=2))


def fp16_ulp_diff(a: torch.Tensor, b: torch.Tensor) -> int:
    """largest distance in fp16 representable steps between two fp16 tensors"""
    ia = a.cpu().view(torch.int16).to(torch.int32)
    ib = b.cpu().view(torch.int16).to(torch.int32)
    # map sign-magnitude to a monotone integer line
    ia = torch.where(ia < 0, -(ia & 0x7fff), ia)
    ib = torch.where(ib < 0, -(ib & 0x7fff), ib)
    return int((ia - ib).abs().max().item()) if ia.numel() else 0


def compare_run(ref: Dict, res: Dict, check_diag: bool = True, tol_merged: Optional[float] = None) -> Dict:
    """Assert parity of one fused run against the oracle; returns a small report."""
    job = res["job"]
    if tol_merged is None:
        tol_merged = TOL_MERGED_FP16B if job.cfg.svd_fp16 else TOL_MERGED
    report = {"params": 0, "code_total": 0, "code_equal": 0, "chigh_total": 0, "chigh_equal": 0, "max_merged_rel": 0.0,
              "flipped_params": [], "dust_params": []}
    flipped, dust = set(), set()
    # same set of parameters got a basis
    assert sorted(ref["bases"].keys()) == sorted(res["bases"].keys())
    # combined masks: bit-exact
    cm = job.combined_masks()
    assert sorted(cm.keys()) == sorted(k for k in ref["combined_masks"] if k in job.shapes)
    for name, m in cm.items():
        assert torch.equal(m.cpu(), ref["combined_masks"][name]), f"combined mask differs: {name}"
    regions = [(name, rb, ref["compressed"][name], "masked") for name, rb in ref["bases"].items()]
    # noise region (svd_include_noise): the same checks on the basis / coefficients of the unmasked rows
    ref_noise = ref.get("bases_noise") or {}
    regions += [(name, rb, ref["compressed_noise"][name], "noise") for name, rb in ref_noise.items()]
    if job.noise:
        for name in ref["bases"]:
            assert res["bases"].meta(name, "noise")["solved"] == (name in ref_noise), f"noise basis presence: {name}"
    for name, rb, ref_comp, region in regions:
        report["params"] += 1
        meta = res["bases"].meta(name, region)
        S_ref = rb["singular_values"].numpy()
        dt, p = res["bases"]._index[name]
        S_new = job._fetch()[dt]["sv" if region == "masked" else "sv_n"][p][: len(S_ref)]
        assert meta["D"] == rb["D"] and meta["N"] == rb["N"]
        assert len(S_ref) == meta["r"]
        if S_ref[0] > 0:
            # Gram route: lambda_j is resolved to ~eps_G * lambda_1, so sigma_j to eps_G * sigma_1^2 / (2 sigma_j);
            # that exceeds TOL_S * sigma_1 only for directions below ~1e-3 sigma_1 (< 1e-6 of the energy)
            # ... the Gram entries themselves are sums of D fp32-rounded products: for a handful of rows their round-off
            # (6e-8 / sqrt(D) relative) no longer averages below eps_G, hence the 2e-8 / sqrt(D) term (D = 16: 5e-9;
            # seen 2.2e-9 in 900 random configurations; irrelevant from D = 100 on)
            tol_gram = max(TOL_S_GRAM, 2e-8 / np.sqrt(max(rb["D"], 1)))
            tol_j = np.maximum(TOL_S * S_ref[0], tol_gram * S_ref[0] ** 2 / np.maximum(S_ref, 1e-30))
            null = S_ref <= 1e-5 * S_ref[0]                     # numerically-null direction: round-off in both
            tol_j[null] = 1e-4 * S_ref[0]
            assert (np.abs(S_new - S_ref) <= tol_j + 1e-30).all(), f"singular values differ: {name}"
        assert meta["k"] == rb["k"], f"rank differs for {name}: {meta['k']} vs {rb['k']}"
        assert abs(meta["energy_retained"] - rb["energy_retained"]) <= 1e-5
        # A low-energy block of <= 2 coefficients one of which belongs to the numerically-null direction of
        # the centred task matrix: whether the reference's stage >= 1 scale is finite or inf/NaN is decided
        # by LAPACK round-off dust, so NaN-ness of this parameter is not reproducible.
        if job.cfg.svd_center and 0 < meta["r"] - meta["k"] <= 2 and meta["r"] == meta["N"]:
            dust.add(name)
        # A low-energy block of exactly 2 coefficients with more than one RTVQ stage: after stage 0 the two residuals
        # are mathematically EQUAL (codes 0 and 2^b-1: x - deq = (scale*lo - round(scale*lo)) / scale for both), so
        # the stage-1 range is pure round-off of the coefficients -- exactly 0 (scale = inf -> NaN, rtvq.py:17) or a
        # few ulps -- in the reference as well.  Which one is not reproducible from coefficients that differ by 1e-7.
        if job.stages > 1 and meta["r"] - meta["k"] == 2:
            dust.add(name)
        # A 2-element low block quantises to the codes {0, 2^b - 1}: both dequantise almost exactly and the whole error
        # is the zero-point rounding residual frac(scale * lo) / scale, which moves by percents when the range
        # (a difference of two close coefficients) changes by 1e-5 -- compared at the looser "flipped" tolerances.
        if meta["r"] - meta["k"] == 2:
            flipped.add(name)
        # Generally: every RTVQ stage maps the two extremes of its input to the end codes, whose residuals are
        # round-off; the numerically-null coefficient of a centred full-rank block is round-off from the start.  Once
        # all n_low elements are round-off, the next stage's range is 0 or a few ulps (inf / NaN or a huge scale).
        n_low_eff = (meta["r"] - meta["k"]) - (1 if (job.cfg.svd_center and meta["r"] == meta["N"]) else 0)
        if job.stages > 1 and 0 < meta["r"] - meta["k"] and n_low_eff <= 2 * (job.stages - 1):
            dust.add(name)
        # Each RTVQ stage removes `bits` bits of the residual; once bits * (stages - 1) exceeds fp32's 24-bit
        # significand the later residuals are round-off (or exactly zero -> scale = inf -> NaN, rtvq.py:17): not
        # reproducible from coefficients that differ in the last place.
        if meta["r"] - meta["k"] > 0 and job.bits * (job.stages - 1) >= 20:
            dust.add(name)
        # coefficients / codes
        comp_new = res["compressed"][name]
        raw = res["compressed"].raw_coefficients(name, region)
        k = rb["k"]
        for ti, task in enumerate(job.tasks):
            if task not in ref_comp or ref_comp[task] is None:
                continue
            rc = ref_comp[task]
            nc = comp_new[task]["masked" if region == "masked" else "unmasked"]
            has_raw = "c_high_fp32" in rc
            c_ref = torch.cat([rc["c_high_fp32"], rc["c_low_fp32"]]).numpy() if has_raw else np.zeros(0, np.float32)
            finite = np.isfinite(c_ref).all()
            if has_raw and finite and len(c_ref):
                scale = np.abs(c_ref).max()
                # the oracle projects on fp16(U) (cli.py:355-361 before compress.py:18-19); the closed form
                # Sigma V^T differs from that by the fp16 rounding noise of U averaged over D rows
                tol_c = max(20 * TOL_COEF, 3e-3 / np.sqrt(max(meta["D"], 1))) if job.cfg.svd_fp16 else 20 * TOL_COEF
                assert np.abs(raw[ti, : len(c_ref)] - c_ref).max() <= tol_c * scale + 1e-12, \
                    f"coefficients differ: {name}/{task}"
            report["chigh_total"] += k
            eq_h = int((nc["c_high_fp16"].view(torch.int16) == rc["c_high_fp16"].view(torch.int16)).sum())
            report["chigh_equal"] += eq_h
            if eq_h != k:
                flipped.add(name)
            # fp16 high block: within one fp16 ulp of the element, or -- for an element much smaller than the
            # vector's scale, whose own ulp is below the coefficient noise -- within the raw-coefficient tolerance
            a16, b16 = nc["c_high_fp16"].float(), rc["c_high_fp16"].float()
            ulp = torch.maximum(b16.abs() * 2.0 ** -10, torch.tensor(2.0 ** -24))
            tol_h = (max(20 * TOL_COEF, 3e-3 / np.sqrt(max(meta["D"], 1))) if job.cfg.svd_fp16 else 20 * TOL_COEF)
            scale_h = float(b16.abs().max()) if b16.numel() else 0.0
            okh = (a16 - b16).abs() <= torch.maximum(ulp, torch.tensor(tol_h * scale_h))
            assert bool(okh.all()) or not torch.isfinite(b16).all(), f"c_high differs: {name}/{task}"
            pr, pn = rc["c_low_quant"]["payloads"], nc["c_low_quant"]["payloads"]
            assert len(pr) == len(pn)
            for a, b in zip(pr, pn):
                report["code_total"] += a["quantized"].numel()
                eq_c = int((a["quantized"] == b["quantized"]).sum())
                report["code_equal"] += eq_c
                if eq_c != a["quantized"].numel():
                    flipped.add(name)
    # merged weights
    for name, m_ref in ref["merged_state_dict"].items():
        m_new = res["merged_state_dict"][name]
        assert m_new.shape == m_ref.shape and m_new.dtype == m_ref.dtype, name
        if name in dust and not nan_positions_equal(m_new, m_ref):
            continue
        assert nan_positions_equal(m_new, m_ref), f"NaN positions differ: {name}"
        fin = torch.isfinite(m_ref)
        if name in ref["merged_deltas"]:
            d_ref = (m_ref - _base_of(ref, name)).double()[fin]
            d_new = (m_new.cpu() - _base_of(ref, name)).double()[fin]
            den = d_ref.norm().item()
            err = (d_new - d_ref).norm().item() / den if den > 0 else (d_new - d_ref).norm().item()
            report["max_merged_rel"] = max(report["max_merged_rel"], err)
            # a flipped code moves one coefficient by one quantiser step of its stage: scale the bound with the step
            tol_flip = max(TOL_MERGED_FLIP, 0.5 / float((1 << job.bits) - 1) ** job.stages)
            tol = tol_flip if name in flipped else tol_merged
            if job.cfg.svd_fp16 and name in ref["bases"]:
                # fp16 bases over a handful of rows: a single basis entry landing on the other side of its fp16 rounding
                # boundary (5e-4 of that entry) is a visible share of the whole delta: 8e-4 / sqrt(Dm) (Dm = 16: 2e-4;
                # equals TOL_MERGED_FP16B from 64 rows on; worst seen in 900 random configurations: 1.06e-4 at 16 rows)
                tol = max(tol, 8e-4 / np.sqrt(max(ref["bases"][name]["D"], 1)))
            assert err <= tol, f"merged delta differs for {name}: rel L2 {err:.3e} (tol {tol:.0e})"
        else:
            assert torch.equal(m_new.cpu(), m_ref), f"untouched parameter changed: {name}"
    report["flipped_params"] = sorted(flipped)
    report["dust_params"] = sorted(dust)
    import os
    RATES.append({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "params": report["params"],
                  "n_tasks": job.N, "fp16_bases": bool(job.cfg.svd_fp16), "bits": job.bits, "stages": job.stages,
                  "projection": job.projection, "c_high_equal": report["chigh_equal"],
                  "c_high_total": report["chigh_total"], "codes_equal": report["code_equal"],
                  "codes_total": report["code_total"], "max_merged_rel": report["max_merged_rel"],
                  "flipped_params": len(flipped), "dust_params": len(dust)})
    flipped = flipped | dust
    if check_diag and ref["diagnostics"].get("per_parameter") is not None and "per_parameter" in res["diagnostics"]:
        compare_diagnostics(ref["diagnostics"], res["diagnostics"], flipped=flipped, dust=dust,
                            max_abs_tol=0.2 if job.cfg.svd_fp16 else 0.0, fp16_bases=bool(job.cfg.svd_fp16),
                            flipped_tol=5e-2 if job.bits >= 4 else (0.3 if job.bits == 3 else 1.0))
    return report


def _base_of(ref, name):
    return ref["_base"][name].float()


# observed relative deviation of max_absolute_error (fp16 bases, parameters whose stored artifacts equal the oracle's)
MAX_ABS_DEV = {"worst": 0.0, "n": 0}


def compare_diagnostics(d_ref: Dict, d_new: Dict, tol_exact: float = TOL_DIAG, flipped=(), dust=(),
                        max_abs_tol: float = 0.0, fp16_bases: bool = False, flipped_tol: float = 5e-2):
    """Diagnostics floats agree to TOL_DIAG for every parameter whose stored artifacts (fp16 c_high, RTVQ
    codes) are identical to the oracle's; a parameter with a one-step flip is itself a different (equally
    valid) quantisation, so its error figures are only compared at the quantisation-noise level."""
    assert sorted(d_ref["per_parameter"]) == sorted(d_new["per_parameter"])
    for name, pr in d_ref["per_parameter"].items():
        pn = d_new["per_parameter"][name]
        # a flipped code is a different (equally valid) quantisation: its error figures agree only at the level of the
        # quantiser step, which is most of the error itself for 1- and 2-bit codes
        tol = flipped_tol if name in flipped else tol_exact
        if pr.get("basis") and pr["basis"]["D"] <= 2 * pr["basis"]["N"]:
            tol = max(tol, 0.25)          # a handful of rows: every fp16 rounding of a basis entry shows in the figures
        assert pn["original_shape"] == pr["original_shape"]
        assert int(pn["masked_size"]) == int(pr["masked_size"]) and int(pn["unmasked_size"]) == int(pr["unmasked_size"])
        assert pn["basis"]["k"] == pr["basis"]["k"] and pn["basis"]["D"] == pr["basis"]["D"]
        assert pn["compression_ratios"] == pr["compression_ratios"], name
        if name in dust:
            continue                      # NaN-ness of its error figures is decided by round-off (see compare_run)
        for task, er in pr["reconstruction_errors"].items():
            en = pn["reconstruction_errors"][task]
            for key, v in er.items():
                w = en[key]
                if np.isnan(v):
                    assert np.isnan(w), f"{name}/{task}/{key}: expected NaN"
                else:
                    # error figures are norms of (orig - rec), a difference of fp32 quantities each carrying a
                    # few 1e-7 * original_norm of round-off (basis row, coefficient, contraction): absolute floor
                    # of 5e-6 * original_norm (5e-6 for the ratio)
                    # ... with fp16 bases the floor is the basis' fp16 rounding noise (individual entries of U land on
                    # the other side of a rounding boundary): ~2e-5 of the original norm, more for tiny parameters
                    # where single entries matter (8e-4 / sqrt(D), the same scaling as the merged weights; worst seen in 1100
                    # random configurations: 6.0e-4 / sqrt(D) at D = 55; equals the 2e-5 floor from 1600 rows on)
                    floor_rel = max(2e-5, 8e-4 / np.sqrt(max(pr["basis"]["D"], 1))) if fp16_bases else 5e-6
                    floor = floor_rel * (er["original_norm"] if key != "relative_error" else 1.0)
                    # the maximum over elements of a parameter with a flipped code is set by that one code's step
                    # ... and with fp16 bases it can be an extreme value of the basis' fp16 rounding noise
                    tol_k = tol
                    if key == "max_absolute_error":
                        tol_k = 0.5 if name in flipped else max(tol, max_abs_tol)
                        if name not in flipped and fp16_bases and abs(v) > floor:
                            MAX_ABS_DEV["worst"] = max(MAX_ABS_DEV["worst"], abs(w - v) / abs(v))
                            MAX_ABS_DEV["n"] += 1
                    assert abs(w - v) <= tol_k * abs(v) + floor + 1e-12, f"{name}/{task}/{key}: {w} vs {v}"
    tol = flipped_tol if flipped else tol_exact
    for key, v in d_ref["summary"].items():
        if dust and key == "average_reconstruction_error":
            continue
        w = d_new["summary"][key]
        if isinstance(v, float) and np.isnan(v):
            assert np.isnan(w)
        else:
            assert abs(w - v) <= tol * abs(v) + 5e-6, f"summary {key}: {w} vs {v}"


def golden_as_reference(case: Dict) -> Dict:
    """Re-shape one golden pipeline case (outputs of the REAL reference) into the oracle's result layout."""
    base = case["base"]
    bases = {}
    for p, b in case["bases"].items():
        d = dict(b["masked"])
        d["Vh"] = case["Vh"][p]
        bases[p] = d
    compressed = {p: {t: (a["masked"] if a.get("masked") is not None else None) for t, a in per.items()}
                  for p, per in case["compressed"].items()}
    bases_noise, compressed_noise = {}, {}
    for p, b in case["bases"].items():
        if b.get("noise") is not None:
            bases_noise[p] = dict(b["noise"], Vh=case["Vh_noise"][p])
            compressed_noise[p] = {t: a["unmasked"] for t, a in case["compressed"][p].items()
                                   if a.get("unmasked") is not None}
    merged = case["merged_state_dict"]
    deltas = {p: merged[p] - base[p] for p in case["bases"]}
    diag = case["diagnostics"]
    return {"merged_state_dict": merged, "merged_deltas": deltas, "bases": bases, "compressed": compressed,
            "bases_noise": bases_noise, "compressed_noise": compressed_noise,
            "combined_masks": case["combined_masks"], "diagnostics": diag, "weights": diag["task_weights"],
            "cluster_assignments": diag.get("cluster_assignments"), "_base": base}
