"""The per-parameter solve core (csrc/k2_core.h) compiled for the HOST (test-only build) against the
oracle and the golden vectors: the same source the GPU runs, checked where there is no GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import svd_hybrid_ref as R
from tests.hostcheck import k2host

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fp16_conversion_matches_ieee_rne():
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s
                         for s in (1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1, 100, 1e4, 1e5)] +
                        [np.array([0, -0.0, 65504, 65519.99, 65520, 1e30, np.inf, -np.inf, 2.0 ** -25, 2.0 ** -24,
                                   2.0 ** -14, 6.1e-5, 2.0 ** -25 * 1.0000001], np.float32)])
    with np.errstate(over="ignore"):
        ref = xs.astype(np.float16).view(np.uint16)
    for x, r in zip(xs, ref):
        assert k2host.f32_to_f16(float(x)) == int(r), x
    hs = np.arange(0, 65536, 3, dtype=np.uint16)
    fr = hs.view(np.float16).astype(np.float32)
    for h, f in zip(hs, fr):
        if not np.isnan(f):
            assert np.float32(k2host.f16_to_f32(int(h))).tobytes() == np.float32(f).tobytes()


def test_select_rank_matches_reference_golden():
    bad = 0
    cases = torch.load(os.path.join(GOLD, "rank_golden.pt"), weights_only=False)
    for c in cases:
        k, er = k2host.select_rank(c["S"].numpy(), c["thr"], c["max_rank"], c.get("min_rank", 1))
        near_tie = bool(((c["cum"] - np.float32(c["thr"])).abs() < 2e-7).any())
        if k != c["k"] and not near_tie:
            bad += 1
        if k == c["k"]:
            assert abs(er - c["cum"][k - 1].item()) <= 2e-7
    assert bad == 0


def test_rtvq_short_bit_exact_vs_reference_golden():
    gold = torch.load(os.path.join(GOLD, "rtvq_golden.pt"), weights_only=False)["rtvq"]
    n_checked = 0
    for case in gold:
        x = case["x"]
        if x.numel() == 0 or x.numel() > 32:
            continue
        codes, sc, zp, rn, deq = k2host.rtvq_short(x.numpy(), case["bits"], case["stages"])
        for s, p in enumerate(case["payloads"]):
            assert np.array_equal(codes[s], p["quantized"].numpy())
            for a, b in ((sc[s], p["scale"].item()), (zp[s], p["zero_point"].item())):
                assert (np.isnan(a) and np.isnan(b)) or np.float32(a).tobytes() == np.float32(b).tobytes()
        g = case["dequantized"].numpy()
        assert np.array_equal(np.isnan(deq), np.isnan(g)) and np.array_equal(deq[~np.isnan(g)], g[~np.isnan(g)])
        n_checked += 1
    assert n_checked >= 8


@pytest.mark.parametrize("name", ["union_uniform", "majority_performance_3stage", "nomask_fp32_nocenter"])
def test_solve_against_reference_golden(name):
    """Gram (fp64, numpy) of the golden inputs -> host-compiled solve -> reference's k, S, fp16 c_high, codes."""
    case = torch.load(os.path.join(GOLD, "pipeline_golden.pt"), weights_only=False)[name]
    tasks, cfgd = case["tasks"], case["config"]
    N = len(tasks)
    center = cfgd.get("svd_center", True)
    stages, bits = cfgd.get("svd_rtvq_stages", 2), cfgd.get("svd_low_bits", 4)
    order = np.asarray(sorted(range(N), key=lambda i: tasks[i]), np.int32)
    w = case["diagnostics"]["task_weights"]
    weights = np.asarray([w[t] for t in tasks], np.float64)
    codes_eq = codes_tot = 0
    for p, gb in case["bases"].items():
        g = gb["masked"]
        m = case["combined_masks"].get(p)
        cols = []
        for t in tasks:
            d = (case["finetuned"][t][p] - case["base"][p]).flatten()
            cols.append((d[m.flatten()] if m is not None else d).double().numpy())
        T = np.stack(cols, 1)
        o = k2host.solve(T.T @ T, T.shape[0], center=center, thr=cfgd["svd_energy_threshold"], max_rank=64, bits=bits,
                         stages=stages, has_mask=m is not None, weights=weights, avg_order=order,
                         sign_ref=case["Vh"][p].double().numpy())
        assert o["info"][0] == 0 and o["info"][3] == g["k"], p
        S = g["singular_values"].numpy()
        assert np.abs(o["sv"][: len(S)] - S).max() <= 2e-6 * S[0]
        assert abs(o["scal"][0] - g["energy_retained"]) <= 1e-5
        k, r = g["k"], len(S)
        for ti, t in enumerate(tasks):
            art = case["compressed"][p][t]["masked"]
            ch_ref = art["c_high_fp16"].view(torch.int16).numpy().astype(np.int32)
            ch_new = o["chigh"][ti, :k].view(np.int16).astype(np.int32)
            assert np.abs(ch_ref - ch_new).max() <= 1, (p, t)
            for s, pay in enumerate(art["c_low_quant"]["payloads"]):
                codes_tot += r - k
                codes_eq += int((pay["quantized"].numpy() == o["codes"][ti, s, : r - k]).sum())
    assert codes_eq >= 0.9 * codes_tot, (codes_eq, codes_tot)


def test_gating_and_degenerate_inputs():
    G = np.eye(4)
    assert k2host.solve(G, 5, has_mask=True, min_mask_size=10)["info"][0] == 1      # mask below svd_min_mask_size
    assert k2host.solve(G, 0)["info"][0] == 2                                        # nothing to decompose
    assert k2host.solve(G, 5, has_mask=False)["info"][0] == 0                        # no mask: no size gate
    z = k2host.solve(np.zeros((4, 4)), 100)                                          # a parameter no task changed
    assert z["info"][3] == 1 and z["info"][4] == 0                                   # k = 1 via the flat-spectrum branch
    assert np.isnan(z["cbar"][1:4]).all() and np.isnan(z["scal"][1])                 # constant low block -> NaN (rtvq.py:17)
    one = k2host.solve(np.eye(8) * 2.0, 1000, thr=0.95, center=True)                 # equal spectrum, centred: k = N-1
    assert one["info"][3] == 7 and np.isnan(one["scal"][1])                          # 1-element low block -> NaN
