"""Reference-shaped views over the packed outputs of a MergeJob.

The reference returns nested dicts of thousands of tiny tensors (cli.py:773-778; layouts in
basis.py:398-407, compress.py:53-56, rtvq.py:69-75,120-126, diagnostics.py:147-231,261-319).
Building them eagerly would cost more host time than the whole GPU merge, so they are
materialised per parameter on access from a handful of small device->host copies.
"""
from __future__ import annotations

from collections import OrderedDict
from collections.abc import Mapping
from typing import Dict, Tuple

import numpy as np
import torch


def estimate_ratio(n: int, bits: int, stages: int) -> float:
    """estimate_compression_ratio (src/svd_hybrid/rtvq.py:142-161) for an n-element low block."""
    return (n * 4) / max(n * bits / 8 * stages + 8 * stages, 1)


class _LazyParamMap(Mapping):
    def __init__(self, job, index: Dict[str, Tuple[torch.dtype, int]]):
        self._job = job
        self._index = OrderedDict((k, index[k]) for k in sorted(index))
        self._cache: Dict[str, dict] = {}

    def __len__(self):
        return len(self._index)

    def __iter__(self):
        return iter(self._index)

    def __contains__(self, k):
        return k in self._index

    def __getitem__(self, k):
        if k not in self._cache:
            self._cache[k] = self._build(k)
        return self._cache[k]

    def _build(self, k):
        raise NotImplementedError


class LazyCompressed(_LazyParamMap):
    """compressed[param][task] = {"masked": {"c_high_fp16", "c_low_quant"}, "unmasked": same | None}
    (compress.py:77-108; "unmasked" only with svd_include_noise and a non-empty noise region)."""

    def _region(self, f, p, t, sfx):
        job = self._job
        info = f["info" + sfx][p]
        if info[0] != 0:
            return None
        r, k = int(info[2]), int(info[3])
        n_low = r - k
        c_high = torch.from_numpy(f["chigh" + sfx][p, t, :k].copy().view(np.float16))
        payloads = []
        if n_low > 0:
            for s in range(job.stages):
                payloads.append({
                    "stage": s,
                    "quantized": torch.from_numpy(f["codes" + sfx][p, t, s, :n_low].copy()),
                    "scale": torch.tensor(f["qscale" + sfx][p, t, s]),
                    "zero_point": torch.tensor(f["qzp" + sfx][p, t, s]),
                    "residual_norm": float(f["qres" + sfx][p, t, s]),
                })
        return {"c_high_fp16": c_high,
                "c_low_quant": {"payloads": payloads, "num_bits": job.bits, "num_stages": job.stages,
                                "original_shape": torch.Size([n_low]), "original_dtype": "torch.float32"}}

    def _build(self, name):
        job = self._job
        dt, p = self._index[name]
        f = job._fetch()[dt]
        present = int(job.groups[dt].host["present"][p])
        out = OrderedDict()
        for t, task in enumerate(job.tasks):
            if not (present >> t) & 1:
                continue
            out[task] = {"masked": self._region(f, p, t, ""),
                         "unmasked": self._region(f, p, t, "_n") if job.noise else None}
        return out

    def raw_coefficients(self, name, region: str = "masked") -> np.ndarray:
        """coef[t][j] before the fp16 / RTVQ round trip (first r columns valid)."""
        dt, p = self._index[name]
        return self._job._fetch()[dt]["coef" if region == "masked" else "coef_n"][p]


class LazyBases(_LazyParamMap):
    """bases[param] = {"masked": {U_high, U_low, singular_values, k, mean, energy_retained, D, N}, "noise": None}.
    U_high / U_low / mean are written by the K5 kernel on first access."""

    def meta(self, name, region: str = "masked") -> Dict:
        job = self._job
        dt, p = self._index[name]
        f = job._fetch()[dt]
        sfx = "" if region == "masked" else "_n"
        info = f["info" + sfx][p]
        return {"k": int(info[3]), "r": int(info[2]), "D": int(f["dm" + sfx][p]), "N": int(info[1]),
                "energy_retained": float(f["scal" + sfx][p, 0]), "solved": int(info[0]) == 0}

    def _region(self, name, region):
        job = self._job
        dt, p = self._index[name]
        f = job._fetch()[dt]
        m = self.meta(name, region)
        if not m["solved"]:
            return None
        sfx = "" if region == "masked" else "_n"
        uh, ul, mn = job.basis_tensors(dt, p, region)
        sv = torch.from_numpy(f["sv" + sfx][p, : m["r"]].copy()).to(job.device)
        return {"U_high": uh, "U_low": ul, "singular_values": sv, "k": m["k"], "mean": mn,
                "energy_retained": m["energy_retained"], "D": m["D"], "N": m["N"]}

    def _build(self, name):
        job = self._job
        job._materialize_bases()
        return {"masked": self._region(name, "masked"),
                "noise": self._region(name, "noise") if job.noise else None}

    def right_vectors(self, name, region: str = "masked") -> np.ndarray:
        """V[t][j] (fp64), the right singular vectors the coefficients were formed from."""
        dt, p = self._index[name]
        return self._job._fetch()[dt]["V" if region == "masked" else "V_n"][p]


_ERR_KEYS = ("absolute_error", "relative_error", "max_absolute_error", "mean_absolute_error", "original_norm",
             "reconstructed_norm")


REC_HEAD = 16      # status, n_active, r, k, dm, has_mask, present, numel, energy_retained, ndim, shape[0..5]
REC_MAX_DIMS = 6


def record_width(n_tasks: int) -> int:
    return REC_HEAD + 6 * n_tasks


def pack_records(job, index) -> Tuple[list, np.ndarray]:
    """Per-parameter records of the parameters that received a basis, as one fp64 matrix (every field is an
    int32, a float32 or an fp64 value, so fp64 holds it exactly): what a rank contributes to the diagnostics
    gather of a parameter-sharded merge (ONE flat all-gather instead of pickled dicts), and what
    diagnostics_from_records needs.  -> (names in sorted order, [P, record_width(N)])."""
    fetched = job._fetch()
    names = sorted(index)
    N = job.N
    rec = np.zeros((len(names), record_width(N)), np.float64)
    for i, name in enumerate(names):
        dt, p = index[name]
        g, f = job.groups[dt], fetched[dt]
        info = f["info"][p]
        shape = list(g.shapes[p])
        if len(shape) > REC_MAX_DIMS:
            raise ValueError(f"{name}: more than {REC_MAX_DIMS} dimensions")
        rec[i, :4] = info[:4]
        rec[i, 4] = int(f["dm"][p])
        rec[i, 5] = int(g.host["has_mask"][p])
        rec[i, 6] = int(g.host["present"][p])
        rec[i, 7] = int(g.numel[p])
        rec[i, 8] = float(f["scal"][p, 0])
        rec[i, 9] = len(shape)
        rec[i, 10: 10 + len(shape)] = shape
        if job.want_diag:
            rec[i, REC_HEAD:] = f["diag_out"][p].reshape(-1)
    return names, rec


def diagnostics_from_records(cfg, tasks, bits: int, stages: int, names, rec: np.ndarray, want_diag: bool = True) -> Dict:
    """diagnostics.py:234-321 schema from per-parameter records (pack_records)."""
    out = {"config": {"svd_energy_threshold": cfg.svd_energy_threshold, "svd_max_rank": cfg.svd_max_rank,
                      "svd_low_bits": cfg.svd_low_bits, "svd_rtvq_stages": cfg.svd_rtvq_stages,
                      "svd_mask_strategy": cfg.svd_mask_strategy, "svd_weighting": cfg.svd_weighting},
           "per_parameter": {}, "summary": {}}
    N = len(tasks)
    ranks, energy, errs, ratios = [], [], [], []
    order = sorted(range(len(names)), key=lambda i: names[i])
    for i in order:
        name, row = names[i], rec[i]
        n_active, r, k, dm = int(row[1]), int(row[2]), int(row[3]), int(row[4])
        has_mask, present, numel = bool(row[5]), int(row[6]), int(row[7])
        shape = [int(x) for x in row[10: 10 + int(row[9])]]
        # the reference takes the shape from the FIRST task's vector and leaves None when that task lacks the
        # parameter (diagnostics.py:161-165)
        d = {"param_name": name, "original_shape": shape if (present & 1) else None,
             "masked_size": dm if has_mask else np.prod(shape),
             "unmasked_size": (numel - dm) if has_mask else 0,
             "reconstruction_errors": {}, "compression_ratios": {},
             "basis": {"k": k, "D": dm, "N": n_active, "energy_retained": float(row[8])}}
        rel = []
        ratio = estimate_ratio(r - k, bits, stages)
        dg = row[REC_HEAD:].reshape(N, 6)
        for t, task in enumerate(tasks):
            if not (present >> t) & 1:
                continue
            d["reconstruction_errors"][task] = {key: float(dg[t, j]) for j, key in enumerate(_ERR_KEYS)}
            d["compression_ratios"][task] = ratio
            rel.append(float(dg[t, 1]))
        if rel:
            d["mean_relative_error"] = float(np.mean(rel))
            d["std_relative_error"] = float(np.std(rel))
            d["max_relative_error"] = float(np.max(rel))
            d["min_relative_error"] = float(np.min(rel))
            errs.append(d["mean_relative_error"])
        ranks.append(k)
        energy.append(d["basis"]["energy_retained"])
        if d["compression_ratios"]:
            ratios.append(np.mean(list(d["compression_ratios"].values())))
        out["per_parameter"][name] = d
    out["summary"] = {"num_parameters": len(out["per_parameter"]),
                      "average_rank": float(np.mean(ranks)) if ranks else 0,
                      "std_rank": float(np.std(ranks)) if ranks else 0,
                      "average_energy_retained": float(np.mean(energy)) if energy else 0,
                      "average_reconstruction_error": float(np.mean(errs)) if errs else 0,
                      "average_compression_ratio": float(np.mean(ratios)) if ratios else 0}
    return out


def build_diagnostics(job, index) -> Dict:
    """diagnostics.py:234-321 schema from the fused K3 reductions."""
    names, rec = pack_records(job, index)
    return diagnostics_from_records(job.cfg, job.tasks, job.bits, job.stages, names, rec)
