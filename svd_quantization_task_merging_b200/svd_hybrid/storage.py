"""Artifact persistence.  Mirror of src/svd_hybrid/storage.py:52-409 -- the on-disk layout is part of
the contract:

    artifact_dir/basis/<param>.pt    {"masked": {U_high, U_low, singular_values, k, mean, energy_retained, D, N}[, "noise"]}
    artifact_dir/coeffs/<param>.pt   {task: {"masked": {"c_high_fp16", "c_low_quant"}[, "unmasked"]}}
    artifact_dir/diagnostics.json, artifact_dir/config.json (= asdict(config))
    output_dir/merged_state_dict.pt
File names replace "/" and "\\" in the parameter name by "_"; tensors are stored on the CPU; RTVQ codes
stay one uint8 per code (the reference never bit-packs)."""
import json
import os
from dataclasses import asdict
from typing import Any, Dict

import torch


def _safe(name: str) -> str:
    return name.replace("/", "_").replace("\\", "_")


def _region_to_cpu(reg: Dict) -> Dict:
    return {"U_high": reg["U_high"].cpu(), "U_low": reg["U_low"].cpu(), "singular_values": reg["singular_values"].cpu(),
            "k": reg["k"], "mean": reg["mean"].cpu() if reg["mean"] is not None else None,
            "energy_retained": reg["energy_retained"], "D": reg["D"], "N": reg["N"]}


def save_basis(basis: Dict, param_name: str, output_dir: str):
    d = os.path.join(output_dir, "basis")
    os.makedirs(d, exist_ok=True)
    data = {}
    for region in ("masked", "noise"):
        if basis.get(region) is not None:
            data[region] = _region_to_cpu(basis[region])
    torch.save(data, os.path.join(d, f"{_safe(param_name)}.pt"))


def load_basis(param_name: str, artifact_dir: str, device: str = "cpu") -> Dict:
    path = os.path.join(artifact_dir, "basis", f"{_safe(param_name)}.pt")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Basis file not found: {path}")
    return torch.load(path, map_location=device, weights_only=False)


def save_compressed_coefficients(compressed: Dict[str, Dict[str, Dict]], output_dir: str):
    d = os.path.join(output_dir, "coeffs")
    os.makedirs(d, exist_ok=True)
    for param_name, per_task in compressed.items():
        data = {}
        for task, art in per_task.items():
            entry = {}
            for region in ("masked", "unmasked"):
                if art.get(region) is not None:
                    entry[region] = {"c_high_fp16": art[region]["c_high_fp16"].cpu(),
                                     "c_low_quant": art[region]["c_low_quant"]}
            data[task] = entry
        torch.save(data, os.path.join(d, f"{_safe(param_name)}.pt"))


def load_compressed_coefficients(param_name: str, artifact_dir: str, device: str = "cpu") -> Dict[str, Dict]:
    path = os.path.join(artifact_dir, "coeffs", f"{_safe(param_name)}.pt")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Coefficients file not found: {path}")
    return torch.load(path, map_location=device, weights_only=False)


def _jsonable(obj: Any) -> Any:
    if isinstance(obj, dict):
        return {k: _jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)) and not isinstance(obj, torch.Size):
        return [_jsonable(v) for v in obj]
    if isinstance(obj, torch.Size):
        return list(obj)
    if isinstance(obj, torch.Tensor):
        return obj.cpu().tolist() if obj.numel() > 1 else obj.item()
    if hasattr(obj, "item"):
        return obj.item()
    return obj


def save_diagnostics(diagnostics: Dict[str, Any], output_dir: str):
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "diagnostics.json"), "w") as f:
        json.dump(_jsonable(diagnostics), f, indent=2)


def load_diagnostics(artifact_dir: str) -> Dict[str, Any]:
    path = os.path.join(artifact_dir, "diagnostics.json")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Diagnostics file not found: {path}")
    with open(path, "r") as f:
        return json.load(f)


def save_config(config, output_dir: str):
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "config.json"), "w") as f:
        json.dump(asdict(config), f, indent=2)


def load_config(artifact_dir: str):
    from .config import SVDHybridConfig
    path = os.path.join(artifact_dir, "config.json")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Config file not found: {path}")
    with open(path, "r") as f:
        return SVDHybridConfig(**json.load(f))


def save_all_artifacts(bases: Dict[str, Dict], compressed: Dict[str, Dict[str, Dict]], diagnostics: Dict[str, Any],
                       config, output_dir: str):
    os.makedirs(output_dir, exist_ok=True)
    for name in bases:
        save_basis(bases[name], name, output_dir)
    save_compressed_coefficients(compressed, output_dir)
    save_diagnostics(diagnostics, output_dir)
    save_config(config, output_dir)


def load_all_artifacts(artifact_dir: str, device: str = "cpu") -> Dict[str, Any]:
    """Parameter list comes from diagnostics["per_parameter"] (storage.py:364)."""
    config = load_config(artifact_dir)
    diagnostics = load_diagnostics(artifact_dir)
    bases, compressed = {}, {}
    for name in diagnostics.get("per_parameter", {}).keys():
        try:
            bases[name] = load_basis(name, artifact_dir, device)
        except FileNotFoundError:
            print(f"Warning: Basis not found for {name}")
        try:
            compressed[name] = load_compressed_coefficients(name, artifact_dir, device)
        except FileNotFoundError:
            print(f"Warning: Coefficients not found for {name}")
    return {"bases": bases, "compressed": compressed, "diagnostics": diagnostics, "config": config}


def save_combined_masks(masks: Dict[str, torch.Tensor], output_dir: str, filename: str = "combined_masks.pt"):
    """Not part of the reference layout (it stores no masks): one extra file next to config.json, bit-packed
    (numpy packbits, little bit order) with the shapes, that lets reload re-merge masked runs."""
    import numpy as np
    os.makedirs(output_dir, exist_ok=True)
    packed = {name: {"shape": list(m.shape),
                     "bits": torch.from_numpy(np.packbits(m.detach().cpu().numpy().reshape(-1), bitorder="little"))}
              for name, m in masks.items()}
    torch.save(packed, os.path.join(output_dir, filename))


def load_combined_masks(artifact_dir: str, device: str = "cpu", filename: str = "combined_masks.pt"
                        ) -> Dict[str, torch.Tensor]:
    """{} when the artifacts hold no masks (unmasked run, or artifacts written by the reference)."""
    import numpy as np
    path = os.path.join(artifact_dir, filename)
    if not os.path.exists(path):
        return {}
    out = {}
    for name, rec in torch.load(path, map_location="cpu", weights_only=False).items():
        n = int(np.prod(rec["shape"])) if len(rec["shape"]) else 1
        bits = np.unpackbits(rec["bits"].numpy(), count=n, bitorder="little").astype(bool)
        out[name] = torch.from_numpy(bits).view(rec["shape"]).to(device)
    return out


def save_merged_model(merged_state_dict: Dict[str, torch.Tensor], output_dir: str, filename: str = "merged_state_dict.pt"):
    os.makedirs(output_dir, exist_ok=True)
    torch.save({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in merged_state_dict.items()},
               os.path.join(output_dir, filename))


def load_merged_model(path: str, device: str = "cpu", filename: str = "merged_state_dict.pt") -> Dict[str, torch.Tensor]:
    """Counterpart of save_merged_model; ``path`` is the file or the directory it was saved into.  Additive: the
    reference's scripts/reload_svd_hybrid.py:18 imports this name, but its storage.py (392-409) only has the saver."""
    if os.path.isdir(path):
        path = os.path.join(path, filename)
    return torch.load(path, map_location=device, weights_only=False)
