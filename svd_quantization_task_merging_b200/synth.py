"""Synthetic checkpoints of the shapes BASELINE.json names (no network: no real checkpoints).

Tensor names/shapes follow open_clip ``model.visual.state_dict()`` for the CLIP
encoders (the reference notebook builds its base state dict that way) and the
HF Llama layout for the capacity config.  Two input families:

* ``throughput``: base ~ N(0, 0.02^2), deltas iid N(0, 0.01^2).
* ``parity``:     deltas with a decaying spectrum, ``D = (A diag(d^j)) C + 1e-4 E`` (d = 0.6 up to 8 tasks),
  so that the selected rank stays away from the reference's NaN edge
  (a 1-element low-energy block divides by max-min = 0, rtvq.py:17).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch


def clip_vit_shapes(width: int, layers: int, patch: int, image: int, out_dim: int) -> "OrderedDict[str, Tuple[int, ...]]":
    grid = image // patch
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["class_embedding"] = (width,)
    s["positional_embedding"] = (grid * grid + 1, width)
    s["proj"] = (width, out_dim)
    s["conv1.weight"] = (width, 3, patch, patch)
    s["ln_pre.weight"] = (width,)
    s["ln_pre.bias"] = (width,)
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        s[p + "ln_1.weight"] = (width,)
        s[p + "ln_1.bias"] = (width,)
        s[p + "attn.in_proj_weight"] = (3 * width, width)
        s[p + "attn.in_proj_bias"] = (3 * width,)
        s[p + "attn.out_proj.weight"] = (width, width)
        s[p + "attn.out_proj.bias"] = (width,)
        s[p + "ln_2.weight"] = (width,)
        s[p + "ln_2.bias"] = (width,)
        s[p + "mlp.c_fc.weight"] = (4 * width, width)
        s[p + "mlp.c_fc.bias"] = (4 * width,)
        s[p + "mlp.c_proj.weight"] = (width, 4 * width)
        s[p + "mlp.c_proj.bias"] = (width,)
    s["ln_post.weight"] = (width,)
    s["ln_post.bias"] = (width,)
    return s


def llama3_8b_shapes() -> "OrderedDict[str, Tuple[int, ...]]":
    h, inter, vocab, kv = 4096, 14336, 128256, 1024
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["model.embed_tokens.weight"] = (vocab, h)
    for i in range(32):
        p = f"model.layers.{i}."
        s[p + "input_layernorm.weight"] = (h,)
        s[p + "self_attn.q_proj.weight"] = (h, h)
        s[p + "self_attn.k_proj.weight"] = (kv, h)
        s[p + "self_attn.v_proj.weight"] = (kv, h)
        s[p + "self_attn.o_proj.weight"] = (h, h)
        s[p + "post_attention_layernorm.weight"] = (h,)
        s[p + "mlp.gate_proj.weight"] = (inter, h)
        s[p + "mlp.up_proj.weight"] = (inter, h)
        s[p + "mlp.down_proj.weight"] = (h, inter)
    s["model.norm.weight"] = (h,)
    s["lm_head.weight"] = (vocab, h)
    return s


def toy_shapes() -> "OrderedDict[str, Tuple[int, ...]]":
    """The 3-tensor toy model of the reference's integration test (tests/test_integration.py:12-19)."""
    return OrderedDict([("layer1.weight", (50, 50)), ("layer2.weight", (50, 25)), ("layer3.weight", (25, 12))])


MODEL_SHAPES = {
    "ViT-B-32": lambda: clip_vit_shapes(768, 12, 32, 224, 512),
    "ViT-B-16": lambda: clip_vit_shapes(768, 12, 16, 224, 512),
    "ViT-L-14": lambda: clip_vit_shapes(1024, 24, 14, 224, 768),
    "Llama-3-8B": llama3_8b_shapes,
    "toy": toy_shapes,
}

STANDARD_8_TASKS = ["Cars", "DTD", "EuroSAT", "GTSRB", "MNIST", "RESISC45", "SUN397", "SVHN"]


def model_shapes(name: str) -> "OrderedDict[str, Tuple[int, ...]]":
    if name not in MODEL_SHAPES:
        raise ValueError(f"unknown model {name!r}; known: {sorted(MODEL_SHAPES)}")
    return MODEL_SHAPES[name]()


def total_params(shapes) -> int:
    n = 0
    for shp in shapes.values():
        m = 1
        for d in shp:
            m *= d
        n += m
    return n


def task_names(n: int) -> List[str]:
    if n <= len(STANDARD_8_TASKS):
        return STANDARD_8_TASKS[:n]
    return STANDARD_8_TASKS + [f"Task{i:02d}" for i in range(len(STANDARD_8_TASKS), n)]


def _tensor_seed(seed: int, name: str) -> int:
    return (seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFFFFFF


def make_checkpoints(shapes, tasks: List[str], family: str = "throughput", seed: int = 1234,
                     device: str = "cpu", dtype: torch.dtype = torch.float32, per_tensor: bool = False,
                     ) -> Tuple[Dict[str, torch.Tensor], Dict[str, Dict[str, torch.Tensor]]]:
    """-> (base_state_dict, {task: finetuned_state_dict}) of random-init weights of the given shapes.

    per_tensor: every tensor draws from its own stream seeded by (seed, name), so that a rank that generates
    only its shard of the model gets exactly the tensors a rank generating the whole model gets."""
    g = torch.Generator(device=device).manual_seed(seed)
    n = len(tasks)
    base: Dict[str, torch.Tensor] = OrderedDict()
    fts: Dict[str, Dict[str, torch.Tensor]] = OrderedDict((t, OrderedDict()) for t in tasks)
    mix = None
    if family == "parity":
        q, _ = torch.linalg.qr(torch.randn(n, n, generator=g, device=device, dtype=torch.float32))
        # 0.6^j for up to 8 tasks; for more tasks the decay is slowed so that the weakest direction stays at
        # ~0.028 of the strongest (= 0.6^7), clear of the 1e-4 noise floor: near-degenerate singular values
        # make the singular VECTORS ill-defined and parity against LAPACK meaningless
        decay = 0.6 if n <= 8 else 0.028 ** (1.0 / (n - 1))
        mix = q * (decay ** torch.arange(n, device=device, dtype=torch.float32))[None, :]
    for name, shp in shapes.items():
        numel = 1
        for d in shp:
            numel *= d
        if per_tensor:
            g.manual_seed(_tensor_seed(seed, name))
        b = torch.randn(numel, generator=g, device=device, dtype=torch.float32) * 0.02
        if family == "parity":
            c = torch.randn(n, numel, generator=g, device=device, dtype=torch.float32) * 0.01
            e = torch.randn(n, numel, generator=g, device=device, dtype=torch.float32) * 0.01
            delta = mix @ c + 1e-4 * e
        elif family == "throughput":
            delta = torch.randn(n, numel, generator=g, device=device, dtype=torch.float32) * 0.01
        else:
            raise ValueError(f"unknown family {family!r}")
        base[name] = b.view(shp).to(dtype)
        for i, t in enumerate(tasks):
            fts[t][name] = (b + delta[i]).view(shp).to(dtype)
    return base, fts


def make_masks(shapes, tasks: List[str], p: float, seed: int = 4321, device: str = "cpu", per_tensor: bool = False,
               ) -> Dict[str, Dict[str, torch.Tensor]]:
    """Per-task Bernoulli(p) tall masks as torch.bool state dicts (per_tensor: see make_checkpoints)."""
    g = torch.Generator(device=device).manual_seed(seed)
    out: Dict[str, Dict[str, torch.Tensor]] = OrderedDict()
    for i, t in enumerate(tasks):
        out[t] = OrderedDict()
        for name, shp in shapes.items():
            if per_tensor:
                g.manual_seed(_tensor_seed(seed + 7919 * (i + 1), name))
            out[t][name] = torch.rand(shp, generator=g, device=device) < p
    return out


def performance_table(tasks: List[str]) -> Dict[str, float]:
    """acc_t = 0.5 + 0.05 t."""
    return {t: 0.5 + 0.05 * i for i, t in enumerate(tasks)}
