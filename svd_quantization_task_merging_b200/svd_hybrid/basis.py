"""SVD basis construction and energy rank selection.  Mirror of src/svd_hybrid/basis.py:63-496.

The decomposition runs on the GPU as a single-pass Gram reduction (K1) + a small Jacobi eigensolve
(K2); U = T V Sigma^-1 is written by K5.  Singular vectors are defined up to sign (and, for a
numerically zero singular value, up to the choice of a null direction: that column of U is zero
here, round-off noise in LAPACK)."""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


def stack_and_center(vectors: List[torch.Tensor], center: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """T = stack(vectors, dim=1) [D x N]; mean over tasks per coordinate (basis.py:63-113)."""
    if not vectors:
        raise ValueError("Empty vector list")
    T = torch.stack(vectors, dim=1)
    mean = None
    if center:
        mean = T.mean(dim=1, keepdim=True)
        T = T - mean
    return T, mean


def _spectrum_f32(S: np.ndarray):
    """fp32 energy spectrum with the reference's arithmetic (basis.py:147-156): S**2 in fp32, double-
    accumulated cumsum rounded to fp32 per prefix, fp32 division by the fp32 total."""
    e = (S.astype(np.float32) * S.astype(np.float32)).astype(np.float32)
    tot = np.float32(e.astype(np.float64).sum())
    if tot < np.float32(1e-10):
        return np.ones_like(e)
    cum = np.cumsum(e.astype(np.float64)).astype(np.float32)
    return (cum / tot).astype(np.float32)


def compute_energy_spectrum(singular_values: torch.Tensor) -> torch.Tensor:
    S = singular_values.detach().float().cpu().numpy()
    return torch.from_numpy(_spectrum_f32(S)).to(singular_values.device)


def select_rank(singular_values: torch.Tensor, energy_threshold: float = 0.90, max_rank: Optional[int] = None,
                min_rank: int = 1) -> int:
    """k = #(cum_energy < thr) + 1, clamped to [min_rank, max_rank, len(S)] (basis.py:159-213).
    Host logic on <= 32 numbers; the fused path does the same selection inside K2."""
    cum = _spectrum_f32(singular_values.detach().float().cpu().numpy())
    k = int((cum < np.float32(energy_threshold)).sum()) + 1
    k = max(k, min_rank)
    if max_rank is not None:
        k = min(k, max_rank)
    return min(k, len(singular_values))


def compute_svd(matrix: torch.Tensor, full_matrices: bool = False, use_randomized: bool = False,
                random_rank: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Thin SVD of a tall [D x N] matrix, N <= 16 (basis.py:216-249).  Results on the input's device."""
    from ..engine import factor_columns
    if full_matrices:
        raise ValueError("only the thin SVD (full_matrices=False) is part of the SVD-Hybrid path")
    if matrix.ndim != 2:
        raise ValueError("compute_svd expects a [D x N] matrix")
    D, N = matrix.shape
    b = factor_columns([matrix[:, i] for i in range(N)], center=False, energy_threshold=1.0, max_rank=None)
    U = torch.cat([b["U_high"], b["U_low"]], dim=1)
    r = U.shape[1]
    Vh = b["V"][:N, :r].T.contiguous().to(torch.float32)
    dev = matrix.device
    return U.to(dev), b["singular_values"].to(dev), Vh.to(dev)


def construct_basis(deltas: List[torch.Tensor], energy_threshold: float = 0.90, max_rank: Optional[int] = None,
                    center: bool = True, device: str = "cpu", use_randomized: bool = False, verbose: bool = True) -> Dict:
    """basis.py:252-409 -> {U_high, U_low, singular_values, k, mean, energy_retained, D, N} on ``device``."""
    from ..engine import factor_columns
    if not deltas:
        raise ValueError("Empty delta list")
    b = factor_columns([d.flatten() for d in deltas], center=center, energy_threshold=energy_threshold,
                       max_rank=max_rank)
    out = {"U_high": b["U_high"].to(device), "U_low": b["U_low"].to(device),
           "singular_values": b["singular_values"].to(device), "k": b["k"],
           "mean": b["mean"].to(device) if b["mean"] is not None else None,
           "energy_retained": b["energy_retained"], "D": b["D"], "N": b["N"]}
    if verbose:
        print(f"   basis: D={out['D']} N={out['N']} k={out['k']} energy_retained={out['energy_retained']:.4f}")
    return out


def construct_masked_basis(masked_deltas: List[torch.Tensor], unmasked_deltas: Optional[List[torch.Tensor]],
                           energy_threshold: float = 0.90, max_rank: Optional[int] = None, center: bool = True,
                           device: str = "cpu", include_noise: bool = False, verbose: bool = False) -> Dict:
    """basis.py:412-468 -> {"masked": basis | None, "noise": basis | None}."""
    out = {"masked": None, "noise": None}
    if masked_deltas and len(masked_deltas[0]) > 0:
        out["masked"] = construct_basis(masked_deltas, energy_threshold, max_rank, center, device, verbose=verbose)
    if include_noise and unmasked_deltas and len(unmasked_deltas[0]) > 0:
        out["noise"] = construct_basis(unmasked_deltas, energy_threshold, max_rank, center, device, verbose=verbose)
    return out


def compute_energy_statistics(singular_values: torch.Tensor) -> Dict[str, float]:
    energy = singular_values ** 2
    total = energy.sum().item()
    n = len(singular_values)
    top = energy[0].item() if n > 0 else 0
    return {"total_energy": total, "top_singular_value": singular_values[0].item() if n > 0 else 0,
            "top_energy_ratio": (top / total if total > 0 else 0) if n > 0 else 0, "num_components": n}
