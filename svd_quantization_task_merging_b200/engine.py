"""Fused SVD-Hybrid merge on one B200: K1 (task vectors + tall masks + Gram) -> K2 (per-parameter
solve) -> K3 (reconstruct + merge [+ diagnostics]) [-> K5 (materialise bases)].

This is the fast path behind ``run_svd_hybrid_pipeline`` (reference: src/svd_hybrid/cli.py:73-778,
steps 1-9) and the fine-grained operator mirrors in ``svd_hybrid/``.  All arithmetic runs in
libsvdq.so (include/svdq.h); torch is used for device memory, streams and host<->device copies.

Data layout in HBM
  * inputs stay where the caller put them (one device tensor per parameter and state dict) or,
    when they arrive from the host, are staged into flat per-state-dict arenas;
  * per launch group (= all parameters of one dtype) small int64/int32 tables describe the work:
    pointer tables [P][N+1] / [P][N], numel[P], and a tile list (tile -> parameter, local index);
  * a tile is TILE_ELEMS consecutive elements of one parameter; K1 writes one partial Gram and
    one masked count per tile, the combined tall mask goes to a bit-packed buffer (1 bit/element);
  * K2 outputs are SoA arrays with stride N per parameter (see include/svdq.h);
  * merged parameters are written into one flat fp32 arena and handed out as views.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native
from .svd_hybrid.weighting import cluster_omega, compute_weights

TILE_ELEMS = int(os.environ.get("SVDQ_TILE_ELEMS", "16384"))          # elements per tile (multiple of 1024); fixes the reduction order
MAX_STREAM_TASKS = 16      # register-resident Gram (one K1 launch)
MAX_TASKS = 32             # wide path: mask pack + single-pass staged Gram (K8) + runtime-N pass 2 (K6)
EXACT_MAX_NUMEL = 262144   # projection="auto": parameters up to this size are re-projected on the stored basis
_FLOAT_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_ALIGN = {torch.float32: 16, torch.bfloat16: 16, torch.float16: 16}   # 16 B: TMA bulk-copy source alignment


_PINNED: Dict[Tuple[int, torch.dtype], torch.Tensor] = {}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dev(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)


def _aligned_flat(t: torch.Tensor, device, align: int) -> torch.Tensor:
    """Contiguous, aligned, flattened device view/copy of ``t`` (copy only when needed)."""
    t = t.detach()
    if t.device != device:
        t = t.to(device, non_blocking=True)
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % align:
        t = t.clone()
    return t.view(-1)


class PackedStateDict(OrderedDict):
    """A state dict whose tensors are views into ONE flat buffer (``.flat``), element offsets in
    ``.offsets``.  Lets the engine move a whole checkpoint with a single host->device copy."""
    flat: torch.Tensor
    offsets: Dict[str, int]


def pack_state_dict(sd: Mapping[str, torch.Tensor], pin: bool = False, align_elems: int = 64) -> PackedStateDict:
    """Re-home a (host or device) state dict into one flat buffer; all tensors must share a dtype."""
    dtypes = {v.dtype for v in sd.values()}
    if len(dtypes) != 1:
        raise ValueError("pack_state_dict needs a single dtype per state dict")
    dtype = next(iter(dtypes))
    if dtype == torch.bool:
        align_elems = max(align_elems, 128)      # masks: rows stay 16-byte aligned after the host-side bit pack
    offs, n = {}, 0
    for k, v in sd.items():
        offs[k] = n
        n += (v.numel() + align_elems - 1) // align_elems * align_elems
    dev = next(iter(sd.values())).device
    flat = torch.zeros(n, dtype=dtype, device=dev)
    if pin and dev.type == "cpu" and torch.cuda.is_available():
        flat = flat.pin_memory()
    out = PackedStateDict()
    for k, v in sd.items():
        view = flat[offs[k]: offs[k] + v.numel()].view(v.shape)
        view.copy_(v)
        out[k] = view
    out.flat, out.offsets = flat, offs
    return out


def _to_device_state_dict(sd: Mapping[str, torch.Tensor], device, wanted=None) -> Tuple[Mapping[str, torch.Tensor], int]:
    """Host -> device move of a state dict -> (device state dict, bytes copied).

    A packed state dict moves with one copy.  With ``wanted`` (this rank's shard of a parameter-sharded merge)
    only those tensors are copied, into one device arena: neighbouring wanted tensors of a packed state dict
    travel as one copy, the rest one copy each -- nothing of another rank's shard crosses PCIe or occupies HBM."""
    if wanted is None:
        if isinstance(sd, PackedStateDict) and sd.flat.device != device:
            flat = sd.flat.to(device, non_blocking=True)
            out = PackedStateDict((k, flat[o: o + v.numel()].view(v.shape)) for (k, v), o in
                                  zip(sd.items(), sd.offsets.values()))
            out.flat, out.offsets = flat, sd.offsets
            return out, flat.numel() * flat.element_size()
        return sd, sum(v.numel() * v.element_size() for v in sd.values() if torch.is_tensor(v) and v.device != device)
    keys = [k for k in sd.keys() if k in wanted and torch.is_tensor(sd[k])]
    host = [k for k in keys if sd[k].device != device]
    out: Dict[str, torch.Tensor] = {k: sd[k] for k in keys if sd[k].device == device}
    if not host:
        return out, 0
    dtypes = {sd[k].dtype for k in host}
    if len(dtypes) != 1:                                  # mixed dtypes: one copy per tensor
        for k in host:
            out[k] = sd[k].to(device, non_blocking=True)
        return out, sum(sd[k].numel() * sd[k].element_size() for k in host)
    dtype = next(iter(dtypes))
    al = 128 if dtype == torch.bool else 64
    packed = isinstance(sd, PackedStateDict) and sd.flat.dtype == dtype
    if packed:
        host.sort(key=lambda k: sd.offsets[k])
    # runs of tensors that are neighbours in the packed host buffer (same padding rule as pack_state_dict)
    runs, total = [], 0
    for k in host:
        n = sd[k].numel()
        span = (n + al - 1) // al * al
        if packed and runs and runs[-1]["src_end"] == sd.offsets[k]:
            r = runs[-1]
            r["keys"].append((k, total, n))
            r["src_end"] += span
            r["valid"] = (total - r["dst"]) + n
        else:
            runs.append(dict(keys=[(k, total, n)], dst=total, valid=n,
                             src=sd.offsets[k] if packed else None, src_end=(sd.offsets[k] + span) if packed else None))
        total += span
    arena = torch.empty(max(total, 1), dtype=dtype, device=device)
    nbytes = 0
    for r in runs:
        if packed:
            arena[r["dst"]: r["dst"] + r["valid"]].copy_(sd.flat[r["src"]: r["src"] + r["valid"]], non_blocking=True)
        else:
            k, o, n = r["keys"][0]
            arena[o: o + n].copy_(sd[k].detach().reshape(-1), non_blocking=True)
        nbytes += r["valid"] * arena.element_size()
        for k, o, n in r["keys"]:
            out[k] = arena[o: o + n].view(sd[k].shape)
    return out, nbytes


def upload_state_dicts(sds: Sequence[Mapping[str, torch.Tensor]], device, n_buffers: int = 2) -> List[Mapping[str, torch.Tensor]]:
    """Pageable host state dicts (what torch.load yields; reference: load_checkpoint / load_task_vectors,
    src/svd_hybrid/task_vector_loader.py:56-189) -> device, PIPELINED: while the DMA of state dict t runs out of one
    pinned staging buffer, the host packs state dict t+1 into the other.  Each state dict arrives as one flat device
    buffer (PackedStateDict).  State dicts with mixed dtypes, or already on the device, are passed through."""
    device = torch.device(device)
    out: List[Mapping[str, torch.Tensor]] = []
    staging: List[Optional[torch.Tensor]] = [None] * n_buffers
    events: List[Optional[torch.cuda.Event]] = [None] * n_buffers
    slot = 0
    for sd in sds:
        tensors = [(k, v) for k, v in sd.items() if torch.is_tensor(v)]
        dtypes = {v.dtype for _, v in tensors}
        if (len(dtypes) != 1 or len(tensors) != len(sd) or any(v.device.type != "cpu" for _, v in tensors)
                or isinstance(sd, PackedStateDict)):
            out.append(sd)
            continue
        dtype = next(iter(dtypes))
        al = 128 if dtype == torch.bool else 64
        offs, n = {}, 0
        for k, v in tensors:
            offs[k] = n
            n += (v.numel() + al - 1) // al * al
        if events[slot] is not None:
            events[slot].synchronize()                      # the previous upload out of this buffer has finished
        buf = staging[slot]
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        if buf is None or buf.numel() < nbytes:
            buf = staging[slot] = torch.empty(nbytes + nbytes // 16, dtype=torch.uint8).pin_memory()
        host = buf[:nbytes].view(dtype)
        for k, v in tensors:
            host[offs[k]: offs[k] + v.numel()].copy_(v.detach().reshape(-1))
        flat = torch.empty(n, dtype=dtype, device=device)
        flat.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        events[slot] = ev
        slot = (slot + 1) % n_buffers
        d = PackedStateDict((k, flat[offs[k]: offs[k] + v.numel()].view(v.shape)) for k, v in tensors)
        d.flat, d.offsets = flat, offs
        out.append(d)
    for ev in events:
        if ev is not None:
            ev.synchronize()                                # staging buffers are released on return
    return out


@dataclass
class _PackedBits:
    """One task mask as it sits on the device after a host-side bit pack: ``bits`` = ceil(numel/8) bytes."""
    shape: torch.Size
    bits: torch.Tensor


_PIN_STAGING: Dict[str, object] = {"buf": None, "event": None}


def _pinned_staging(nbytes: int) -> torch.Tensor:
    """Process-wide pinned staging buffer for the bit-packed masks (grow-only; a previous upload out of it is
    waited for before it is overwritten)."""
    ev = _PIN_STAGING["event"]
    if ev is not None:
        ev.synchronize()
        _PIN_STAGING["event"] = None
    buf = _PIN_STAGING["buf"]
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes + nbytes // 8, dtype=torch.uint8).pin_memory()
        _PIN_STAGING["buf"] = buf
    return buf


def _upload_mask_bits(task_masks: Sequence[Optional[Mapping[str, torch.Tensor]]], device, wanted=None
                      ) -> Tuple[List[Optional[Dict[str, _PackedBits]]], int]:
    """HOST torch.bool masks -> bits on the host (svdq_host_pack_mask, all host threads; runs while the tensors'
    host->device copies are in flight) -> one upload of numel/8 bytes per task.  A pure transfer encoding: the
    masks are combined on the device (K1) exactly as with byte masks."""
    # host threads for the re-encoding: this process's CPUs, shared with the other ranks of the node under torchrun
    n_threads = max(1, min(16, len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    if wanted is not None:                                    # this rank's shard only
        task_masks = [None if m is None else {k: v for k, v in m.items() if k in wanted} for m in task_masks]
    plan, total = [], 0
    for m in task_masks:
        if m is None:
            plan.append(None)
            continue
        whole = (isinstance(m, PackedStateDict) and m.flat.dtype == torch.bool and m.flat.is_contiguous()
                 and all(o % 128 == 0 for o in m.offsets.values()))
        if whole:                                             # one call for the whole flat buffer
            entry = dict(whole=True, base=total, offs={k: o // 8 for k, o in m.offsets.items()})
            total += (m.flat.numel() + 7) // 8
        else:
            offs = {}
            entry = dict(whole=False, base=total, offs=offs)
            for k, v in m.items():
                offs[k] = total - entry["base"]
                total += ((v.numel() + 7) // 8 + 15) // 16 * 16
        total = (total + 16 + 15) // 16 * 16                  # K1's last partial vector reads a whole 32-bit word
        plan.append(entry)
    stage = _pinned_staging(max(total, 16))
    base_ptr = stage.data_ptr()
    keep, src, cnt, dst = [], [], [], []
    for m, entry in zip(task_masks, plan):
        if m is None:
            continue
        if entry["whole"]:
            src.append(m.flat.data_ptr()); cnt.append(m.flat.numel()); dst.append(base_ptr + entry["base"])
        else:
            for k, v in m.items():
                v = v.detach()
                if v.dtype != torch.bool:
                    v = v.bool()
                v = v.contiguous()
                keep.append(v)
                src.append(v.data_ptr()); cnt.append(v.numel()); dst.append(base_ptr + entry["base"] + entry["offs"][k])
    if src:          # one call for the whole batch: its bytes are split evenly over the host threads
        a_src, a_cnt, a_dst = (np.asarray(x, np.int64) for x in (src, cnt, dst))
        _native.call("svdq_host_pack_mask_batch", a_src.ctypes.data, a_cnt.ctypes.data, a_dst.ctypes.data, len(src), n_threads)
    dev_bits = torch.empty(max(total, 16), dtype=torch.uint8, device=device)
    dev_bits.copy_(stage[: dev_bits.numel()], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _PIN_STAGING["event"] = ev
    out: List[Optional[Dict[str, _PackedBits]]] = []
    for m, entry in zip(task_masks, plan):
        if m is None:
            out.append(None)
            continue
        d = {}
        for k, v in m.items():
            o = entry["base"] + entry["offs"][k]
            d[k] = _PackedBits(v.shape, dev_bits[o: o + (v.numel() + 7) // 8])
        out.append(d)
    return out, total


@dataclass
class _Group:
    """All parameters of one dtype: one K1 / K2 / K3 launch each."""
    dtype: torch.dtype
    names: List[str] = field(default_factory=list)
    shapes: List[torch.Size] = field(default_factory=list)
    numel: List[int] = field(default_factory=list)
    keep: List[torch.Tensor] = field(default_factory=list)     # keeps staged tensors alive
    # device tables / workspaces (filled by MergeJob._build_group)
    t: Dict[str, torch.Tensor] = field(default_factory=dict)
    n_tiles: int = 0
    any_mask: bool = False
    out_off: List[int] = field(default_factory=list)
    host: Dict[str, np.ndarray] = field(default_factory=dict)


class MergeJob:
    """One SVD-Hybrid merge of N fine-tuned state dicts against a base on one GPU.

    prepare (constructor) -> run() (kernel launches, asynchronous) -> results() (reference-shaped
    result dict, built lazily from a handful of small device->host copies).
    """

    def __init__(self, base: Mapping[str, torch.Tensor], finetuned: Mapping[str, Mapping[str, torch.Tensor]],
                 task_masks: Optional[Mapping[str, Optional[Mapping[str, torch.Tensor]]]], config,
                 device: Optional[str] = None, *, sign_ref: Optional[Mapping[str, torch.Tensor]] = None,
                 diagnostics: Optional[bool] = None, materialize_bases: bool = False,
                 performance: Optional[Dict[str, float]] = None,
                 cluster_assignments: Optional[Dict[str, int]] = None, param_filter: Optional[Sequence[str]] = None,
                 tile_elems: int = TILE_ELEMS, cluster_backend: Optional[str] = None,
                 projection: Optional[str] = None, sign_ref_noise: Optional[Mapping[str, torch.Tensor]] = None):
        _native.require_cuda()
        self.cfg = config
        self.device = torch.device(device or "cuda")
        if self.device.type != "cuda":
            raise _native.NativeLibraryError("the SVD-Hybrid merge path runs only on a CUDA device (no CPU fallback)")
        self.tasks: List[str] = list(config.tasks) if getattr(config, "tasks", None) else list(finetuned.keys())
        self.N = len(self.tasks)
        if self.N < 1:
            raise ValueError("Empty delta list")
        if self.N > MAX_TASKS:
            raise ValueError(f"n_tasks={self.N}: at most {MAX_TASKS} task vectors per merge")
        self.wide = self.N > MAX_STREAM_TASKS          # 17..32 tasks: blocked Gram + runtime-N pass 2
        # noise region (svd_include_noise): a second basis over the rows outside the combined mask, reconstructed
        # with svd_noise_shrink into the unmasked positions (cli.py:336-351, basis.py:455-466, merge.py:257-284)
        self.noise = bool(getattr(config, "svd_include_noise", False))
        self.noise_shrink = float(getattr(config, "svd_noise_shrink", 0.5))
        if config.svd_mask_strategy not in _native.STRATEGY_CODE:
            raise ValueError(f"Unknown mask strategy: {config.svd_mask_strategy}")
        self.tile_elems = int(tile_elems)
        if self.wide and self.tile_elems % 3072 != 0:
            self.tile_elems = 12288                      # the staged wide Gram walks tiles in chunks of 512 / 768
        self.want_diag = config.svd_eval_reconstruction if diagnostics is None else bool(diagnostics)
        self.materialize = bool(materialize_bases)
        self.sign_ref = sign_ref
        self.sign_ref_noise = sign_ref_noise
        self.performance = performance
        self.fixed_assignments = cluster_assignments
        self.cluster_mode = config.svd_weighting == "cluster"
        self.cluster_backend = cluster_backend or os.environ.get("SVDQ_CLUSTER_BACKEND", "kmeans")
        # coefficients: "closed" = Sigma V^T for every parameter; "exact" = re-project every parameter on the stored
        # (fp16) basis like the reference does (one more read of the inputs); "auto" = exact for parameters of at
        # most EXACT_MAX_NUMEL elements, where the closed form's 2.4e-4/sqrt(Dm) deviation can flip an fp16 value or
        # an RTVQ code, closed form above (deviation below LAPACK's own round-off).  Only meaningful with fp16 bases.
        self.projection = projection or os.environ.get("SVDQ_PROJECTION", "auto")
        if self.projection not in ("closed", "auto", "exact"):
            raise ValueError(f"projection must be closed | auto | exact, got {self.projection}")
        if not config.svd_fp16:
            self.projection = "closed"
        self.stages = int(config.svd_rtvq_stages)
        self.bits = int(config.svd_low_bits)
        if self.stages > 8:
            raise ValueError("RTVQ stages must be <= 8 in this build")

        # SVDQ_PROFILE=1: host wall-clock of the phases, each closed by a device synchronise (diagnostic runs only)
        self._profile = os.environ.get("SVDQ_PROFILE", "0") == "1"
        self.timing: Dict[str, float] = {}
        import time as _time
        with torch.cuda.device(self.device):
            t0 = _time.perf_counter()
            self._stage_inputs(base, finetuned, task_masks, param_filter)
            if self._profile:
                torch.cuda.synchronize(self.device)
                self.timing["stage_inputs_s"] = _time.perf_counter() - t0
                t0 = _time.perf_counter()
            for g in self.groups.values():
                self._build_group(g)
            if self._profile:
                torch.cuda.synchronize(self.device)
                self.timing["build_tables_s"] = _time.perf_counter() - t0
        self._ran = False
        self._side = None
        self._order_dev = _dev(np.asarray(sorted(range(self.N), key=lambda i: self.tasks[i]), np.int32), self.device)
        self.gram_reduce_hook = None
        self._fetched: Optional[Dict[str, Dict[str, np.ndarray]]] = None
        self.weights: Optional[Dict[str, float]] = None
        self.cluster_assignments: Optional[Dict[str, int]] = None

    # ------------------------------------------------------------------------------------------
    def _stage_inputs(self, base, finetuned, task_masks, param_filter):
        dev = self.device
        wanted = set(param_filter) if param_filter is not None else None
        # with a parameter filter (this rank's shard) only the owned tensors cross PCIe and occupy HBM
        self.h2d_bytes = 0
        base_d, nb = _to_device_state_dict(base, dev, wanted)
        self.h2d_bytes += nb
        fts_d = []
        for t in self.tasks:
            if t in finetuned:
                sd, nb = _to_device_state_dict(finetuned[t], dev, wanted)
                self.h2d_bytes += nb
            else:
                sd = {}
            fts_d.append(sd)
        import time as _time
        _t_issue = _time.perf_counter()
        masks_h = [(task_masks.get(t) if task_masks else None) for t in self.tasks]
        # masks that all sit in host memory (what load_task_masks yields) cross PCIe bit-packed: 1/8 of the bytes
        self.mask_bits = (not self.wide and any(m is not None for m in masks_h) and
                          all(torch.is_tensor(v) and v.device.type == "cpu" for m in masks_h if m is not None
                              for v in m.values()))
        masks_d: List[Optional[Mapping[str, object]]] = []
        if self.mask_bits:
            masks_d, nbytes = _upload_mask_bits(masks_h, dev, wanted)
            self.h2d_bytes += nbytes
            if getattr(self, "_profile", False):
                self.timing["host_mask_pack_s"] = _time.perf_counter() - _t_issue
        else:
            for m in masks_h:
                if m is not None:
                    m, nb = _to_device_state_dict(m, dev, wanted)
                    self.h2d_bytes += nb
                masks_d.append(m)

        self.base_keys = list(base.keys())
        self.base_ref = base
        self.groups: "OrderedDict[torch.dtype, _Group]" = OrderedDict()
        self.passthrough: List[str] = []
        self.filtered_out = set()
        self._tensors: Dict[str, List[Optional[torch.Tensor]]] = {}
        self._masks: Dict[str, List[Optional[torch.Tensor]]] = {}
        self.shapes: Dict[str, torch.Size] = {}
        for name in sorted(base.keys()):
            if wanted is not None and name not in wanted:
                self.filtered_out.add(name)          # another rank's shard: not returned by this job
                continue
            b = base_d[name] if name in base_d else base[name]
            if not torch.is_tensor(b) or b.dtype not in _FLOAT_DTYPES or b.numel() == 0:
                self.passthrough.append(name)
                continue
            row: List[Optional[torch.Tensor]] = [None] * (self.N + 1)
            any_task = False
            for i, sd in enumerate(fts_d):
                f = sd.get(name) if sd else None
                if f is None or f.shape != b.shape:
                    continue
                if f.dtype != b.dtype:
                    raise ValueError(f"{name}: dtype mismatch between base ({b.dtype}) and task {self.tasks[i]} ({f.dtype})")
                row[i + 1] = _aligned_flat(f, dev, _ALIGN[b.dtype])
                any_task = True
            if not any_task:
                self.passthrough.append(name)
                continue
            row[0] = _aligned_flat(b, dev, _ALIGN[b.dtype])
            mrow: List[Optional[torch.Tensor]] = [None] * self.N
            for i, m in enumerate(masks_d):
                mk = m.get(name) if m else None
                if mk is None:
                    continue
                if mk.shape != b.shape:
                    mrow = [None] * self.N       # cli.py:330: a mask of the wrong shape is not used
                    break
                if isinstance(mk, _PackedBits):
                    mrow[i] = mk.bits
                    continue
                if mk.dtype != torch.bool:
                    mk = mk.bool()
                mrow[i] = _aligned_flat(mk, dev, 16)
            g = self.groups.setdefault(b.dtype, _Group(dtype=b.dtype))
            g.names.append(name)
            g.shapes.append(b.shape)
            g.numel.append(b.numel())
            self._tensors[name] = row
            self._masks[name] = mrow
            self.shapes[name] = b.shape

    # ------------------------------------------------------------------------------------------
    def _build_group(self, g: _Group):
        dev, N, S, te = self.device, self.N, self.stages, self.tile_elems
        P = len(g.names)
        numel = np.asarray(g.numel, np.int64)
        tiles_per = (numel + te - 1) // te
        tile_begin = np.zeros(P + 1, np.int64)
        np.cumsum(tiles_per, out=tile_begin[1:])
        n_tiles = int(tile_begin[-1])
        tile_param = np.repeat(np.arange(P, dtype=np.int32), tiles_per)
        tile_local = (np.arange(n_tiles, dtype=np.int64) - np.repeat(tile_begin[:-1], tiles_per)).astype(np.int32)
        tptr = np.zeros((P, N + 1), np.int64)
        mptr = np.zeros((P, N), np.int64)
        present = np.zeros(P, np.uint32)
        has_mask = np.zeros(P, np.uint8)
        pm_off = np.zeros(P, np.int64)
        words = 0
        out_off, out_total = [], 0
        for p, name in enumerate(g.names):
            row, mrow = self._tensors[name], self._masks[name]
            g.keep.extend(x for x in row if x is not None)
            g.keep.extend(x for x in mrow if x is not None)
            for i, x in enumerate(row):
                tptr[p, i] = 0 if x is None else x.data_ptr()
            for i in range(N):
                if row[i + 1] is not None:
                    present[p] |= np.uint32(1 << i)
                if mrow[i] is not None:
                    mptr[p, i] = mrow[i].data_ptr()
                    has_mask[p] = 1
            pm_off[p] = words
            if has_mask[p]:
                words += ((int(numel[p]) + 31) // 32 + 32 + 3) // 4 * 4     # rows start 16-byte aligned (TMA source)
            out_off.append(out_total)
            out_total += (int(numel[p]) + 63) // 64 * 64
        g.any_mask = bool(has_mask.any())
        g.n_tiles, g.out_off = n_tiles, out_off
        g.host = dict(numel=numel, tile_begin=tile_begin, has_mask=has_mask, present=present)
        G = N * (N + 1) // 2
        full = 2 if (self.cluster_mode or self.noise) else 1
        f32, i32, i64, f64, u8 = torch.float32, torch.int32, torch.int64, torch.float64, torch.uint8
        z = lambda *shape, dtype=f32: torch.zeros(*shape, dtype=dtype, device=dev)   # noqa: E731

        def solve_outputs():
            return dict(info=z(P, 8, dtype=i32), sv=z(P, N), scal=z(P, 4), coef=z(P, N, N),
                        chigh=z(P, N, N, dtype=torch.int16), codes=z(P, N, S, N, dtype=u8), qscale=z(P, N, S),
                        qzp=z(P, N, S), qres=z(P, N, S), chat=z(P, N, N), cbar=z(P, N), W=z(P, N, N), gvec=z(P, N),
                        V=z(P, N, N, dtype=f64))
        g.t = dict(
            tptr=_dev(tptr, dev), mptr=_dev(mptr, dev) if g.any_mask else None, numel=_dev(numel, dev),
            tile_param=_dev(tile_param, dev), tile_local=_dev(tile_local, dev), tile_begin=_dev(tile_begin, dev),
            pm_off=_dev(pm_off, dev), has_mask=_dev(has_mask, dev), present=_dev(present.view(np.int32), dev),
            packed=z(max(words, 1), dtype=i32), gram=z(max(n_tiles, 1) * full * G), count=z(max(n_tiles, 1), dtype=i32),
            gram_masked=z(P, N * N, dtype=f64), gram_all=z(P, N * N, dtype=f64) if self.cluster_mode else None,
            dm=z(P, dtype=i64), **solve_outputs(),
            out=torch.empty(out_total, dtype=f32, device=dev),
            diag=z(max(n_tiles, 1) * 4 * N) if self.want_diag else None,
            diag_out=z(P, N, 6, dtype=f64) if self.want_diag else None,
        )
        # noise region: its own Gram, row count and solve outputs (same layout as the masked region's)
        g.tn = None
        if self.noise:
            g.tn = dict(gram=z(P, N * N, dtype=f64), dm=z(P, dtype=i64), no_gate=z(P, dtype=u8), **solve_outputs())
        # exact-projection selection: tiles of the selected parameters
        g.sel = None
        if self.projection != "closed":
            pick = np.ones(P, bool) if self.projection == "exact" else (numel <= EXACT_MAX_NUMEL)
            if pick.any():
                sel_tiles = np.where(pick, tiles_per, 0)
                sel_begin = np.zeros(P + 1, np.int64)
                np.cumsum(sel_tiles, out=sel_begin[1:])
                n_sel = int(sel_begin[-1])
                sp = np.repeat(np.arange(P, dtype=np.int32), sel_tiles)
                sl = (np.arange(n_sel, dtype=np.int64) - np.repeat(sel_begin[:-1], sel_tiles)).astype(np.int32)
                g.sel = dict(n=n_sel, begin=_dev(sel_begin, dev), param=_dev(sp, dev), local=_dev(sl, dev),
                             proj=z(max(n_sel, 1) * N * N))
        if self.wide:
            g.t["dm_scratch"] = z(P, dtype=i64)
        optr = np.asarray([g.t["out"].data_ptr() + 4 * o for o in out_off], np.int64)
        g.t["optr"] = _dev(optr, dev)
        def sign_table(ref):
            if ref is None:
                return None
            sr = np.zeros((P, N, N), np.float64)
            for p, name in enumerate(g.names):
                v = ref.get(name)
                if v is not None:
                    v = np.asarray(v.detach().double().cpu().numpy() if torch.is_tensor(v) else v, np.float64)
                    sr[p, : v.shape[0], : v.shape[1]] = v
            return _dev(sr, dev)

        g.t["sign_ref"] = sign_table(self.sign_ref)
        if g.tn is not None:
            g.tn["sign_ref"] = sign_table(self.sign_ref_noise)

    # ------------------------------------------------------------------------------------------
    def _weights_table(self) -> Tuple[torch.Tensor, torch.Tensor]:
        cfg = self.cfg
        if cfg.svd_weighting == "performance" and self.performance is not None:
            from .svd_hybrid.weighting import compute_performance_weights
            self.weights = compute_performance_weights(
                {t: float(self.performance.get(t, 1.0)) for t in self.tasks}, cfg.svd_weighting_temperature)
        else:
            self.weights = compute_weights(self.tasks, weighting_strategy=cfg.svd_weighting,
                                           performance_file=getattr(cfg, "performance_file", None),
                                           temperature=cfg.svd_weighting_temperature,
                                           cluster_assignments=self.cluster_assignments)
        # one pinned staging buffer and ONE asynchronous upload for the three small tables (task weights, cluster
        # omegas, cluster index per task): this sits on the critical path between the k-means and pass 2
        N = self.N
        if getattr(self, "_wt_host", None) is None:
            self._wt_host = torch.empty(20 * N, dtype=torch.uint8, pin_memory=True)
            self._wt_dev = torch.empty(20 * N, dtype=torch.uint8, device=self.device)
            self._wt_event = None
        if self._wt_event is not None:
            self._wt_event.synchronize()                 # the previous upload has left the staging buffer
        host = self._wt_host.numpy()
        w_h, om_h, cl_h = host[: 8 * N].view(np.float64), host[8 * N: 16 * N].view(np.float64), host[16 * N:].view(np.int32)
        w_h[:] = [self.weights.get(t, 1.0) for t in self.tasks]
        om_h[:] = 0.0
        cl_h[:] = 0
        clustered = bool(self.cluster_mode and self.cluster_assignments)
        if clustered:
            # merge_with_clustering (merge.py:586-626): member weights renormalised inside each cluster PER PARAMETER
            # over the members that have it (K2 average_param), clusters averaged with omega
            om = cluster_omega(self.weights, self.cluster_assignments)
            ids = sorted(om.keys())
            index = {c: i for i, c in enumerate(ids)}
            cl_h[:] = [index[self.cluster_assignments[t]] for t in self.tasks]
            om_h[: len(ids)] = [om[c] for c in ids]
        self._wt_dev.copy_(self._wt_host, non_blocking=True)
        self._wt_event = torch.cuda.Event()
        self._wt_event.record()
        dev = self._wt_dev
        self._cluster_tables = (dev[16 * N:].view(torch.int32), dev[8 * N: 16 * N].view(torch.float64)) if clustered else (None, None)
        return dev[: 8 * N].view(torch.float64), self._order_dev

    def _cluster_begin(self):
        """Whole-model task Gram (K1 by-product) -> pinned host copy on a side stream, so that the
        k-means round trip overlaps the per-parameter solve on the main stream."""
        if self.fixed_assignments is not None:
            return
        # a rank whose shard is empty (more ranks than parameters) still joins the all-reduce with zeros
        tot = None
        for g in self.groups.values():
            part = g.t["gram_all"].sum(dim=0)
            tot = part if tot is None else tot + part
        if tot is None:
            tot = torch.zeros(self.N * self.N, dtype=torch.float64, device=self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._gram_host = torch.empty(self.N * self.N, dtype=torch.float64, pin_memory=True)
        ready = torch.cuda.Event()
        ready.record()
        self._side.wait_event(ready)
        with torch.cuda.stream(self._side):
            if self.gram_reduce_hook is not None:   # multi-GPU: sum the per-rank Grams (sharding.allreduce_gram);
                tot = self.gram_reduce_hook(tot)    # on the side stream, so the solve does not wait for the collective
            self._gram_host.copy_(tot, non_blocking=True)
            self._gram_done = torch.cuda.Event()
            self._gram_done.record()
        self._gram_keep = tot

    def _cluster_end(self):
        """Host k-means on the Gram (clustering.py:198-245)."""
        from .svd_hybrid.clustering import cluster_from_gram
        if self.fixed_assignments is not None:
            self.cluster_assignments = dict(self.fixed_assignments)
            return
        self._gram_done.synchronize()
        gram = self._gram_host.numpy().reshape(self.N, self.N).copy()
        self.whole_model_gram = gram
        self.cluster_assignments = cluster_from_gram(gram, self.tasks, self.cfg.svd_cluster_k, "kmeans",
                                                     backend=self.cluster_backend)

    # ------------------------------------------------------------------------------------------
    def run(self, record_events: bool = False) -> "MergeJob":
        """Launch the whole path on the current stream.  Asynchronous except for the k-means
        round trip of cluster weighting."""
        cfg, N, te = self.cfg, self.N, self.tile_elems
        st = _native.stream_ptr()
        strat = _native.STRATEGY_CODE[cfg.svd_mask_strategy]
        full = 2 if self.noise else (1 if self.cluster_mode else 0)   # second Gram block: complement / all / none
        mms = int(cfg.svd_min_mask_size)
        ev = {}

        def mark(name):
            if record_events:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev[name] = e

        with torch.cuda.device(self.device):
            mark("start")
            if not self.wide:
                for g in self.groups.values():
                    t = g.t
                    _native.call("svdq_tv_mask_gram_bits" if self.mask_bits else "svdq_tv_mask_gram",
                                 _FLOAT_DTYPES[g.dtype], N, strat, full, g.n_tiles, te,
                                 _ptr(t["tptr"]), _ptr(t["mptr"]), _ptr(t["numel"]), _ptr(t["tile_param"]),
                                 _ptr(t["tile_local"]), _ptr(t["pm_off"]), _ptr(t["packed"]), _ptr(t["gram"]),
                                 _ptr(t["count"]), st)
                mark("k1")
                for g in self.groups.values():
                    t = g.t
                    tn = g.tn or {}
                    _native.call("svdq_gram_reduce", N, full, len(g.names), mms, _ptr(t["gram"]), _ptr(t["count"]),
                                 _ptr(t["tile_begin"]), _ptr(t["numel"]), _ptr(t["has_mask"]), _ptr(t["gram_masked"]),
                                 _ptr(t["gram_all"]), _ptr(t["dm"]), _ptr(tn.get("gram")), _ptr(tn.get("dm")), st)
            else:
                # 17..32 tasks: combine the masks once, then the single-pass staged Gram (K8)
                for g in self.groups.values():
                    t = g.t
                    P = len(g.names)
                    _native.call("svdq_mask_pack", N, strat, g.n_tiles, te, _ptr(t["mptr"]), _ptr(t["numel"]),
                                 _ptr(t["tile_param"]), _ptr(t["tile_local"]), _ptr(t["pm_off"]), _ptr(t["packed"]),
                                 _ptr(t["count"]), st)
                    # one staged launch per Gram: rows inside the mask, then (cluster weighting / noise region)
                    # all rows or the rows outside the mask
                    second = 2 if self.noise else (1 if self.cluster_mode else 0)
                    dst2 = g.tn["gram"] if self.noise else t["gram_all"]
                    launches = [(0, t["gram_masked"], t["dm"])]
                    if second:
                        launches.append((second, dst2, t["dm_scratch"]))
                    for mode, dst, dm_out in launches:
                        _native.call("svdq_gram_staged", _FLOAT_DTYPES[g.dtype], N, mode, g.n_tiles, te,
                                     _ptr(t["tptr"]), _ptr(t["numel"]), _ptr(t["tile_param"]), _ptr(t["tile_local"]),
                                     _ptr(t["pm_off"]), _ptr(t["has_mask"]), _ptr(t["packed"]), _ptr(t["gram"]), st)
                        _native.call("svdq_gram_reduce", N, 0, P, mms, _ptr(t["gram"]), _ptr(t["count"]),
                                     _ptr(t["tile_begin"]), None, None, _ptr(dst), None, _ptr(dm_out), None, None, st)
                    if self.noise:
                        if self.cluster_mode:            # all rows = masked + unmasked
                            torch.add(t["gram_masked"], g.tn["gram"], out=t["gram_all"])
                        # rows of the noise region (same rule as svdq_gram_reduce's dm_noise)
                        on = (t["has_mask"] != 0) & (t["dm"] >= mms) & (t["dm"] > 0)
                        g.tn["dm"].copy_(torch.where(on, t["numel"] - t["dm"], torch.zeros_like(t["dm"])))
                mark("k1")
            max_rank = int(cfg.svd_max_rank) if cfg.svd_max_rank is not None else 0

            def regions(g):
                # (region code, solve inputs, solve outputs): 0 = rows inside the combined mask; 1 = noise region,
                # solved from the complement Gram without the svd_min_mask_size gate (that gate is already folded
                # into dm_noise by svdq_gram_reduce)
                t = g.t
                out = [(0, t["gram_masked"], t["dm"], t["has_mask"], t)]
                if g.tn is not None:
                    out.append((1, g.tn["gram"], g.tn["dm"], g.tn["no_gate"], g.tn))
                return out

            def solve(w_dev, order_dev):
                for g in self.groups.values():
                    for _, gram, dm, gate, o in regions(g):
                        _native.call("svdq_param_solve", N, len(g.names), int(bool(cfg.svd_center)),
                                     float(cfg.svd_energy_threshold), max_rank, mms, self.bits,
                                     self.stages, _ptr(gram), _ptr(dm), _ptr(gate),
                                     _ptr(g.t["present"]), _ptr(w_dev), _ptr(order_dev), _ptr(o["sign_ref"]),
                                     _ptr(o["info"]), _ptr(o["sv"]), _ptr(o["scal"]), _ptr(o["coef"]),
                                     _ptr(o["chigh"]), _ptr(o["codes"]), _ptr(o["qscale"]), _ptr(o["qzp"]),
                                     _ptr(o["qres"]), _ptr(o["chat"]), _ptr(o["cbar"]), _ptr(o["W"]),
                                     _ptr(o["gvec"]), _ptr(o["V"]), st)

            def project_exact():
                # re-project the selected (small) parameters on the stored basis and re-quantise them (K7)
                for g in self.groups.values():
                    if g.sel is None:
                        continue
                    t, sel = g.t, g.sel
                    for reg, _, _, _, o in regions(g):
                        _native.call("svdq_project_exact", _FLOAT_DTYPES[g.dtype], N, int(bool(cfg.svd_fp16)),
                                     int(bool(cfg.svd_center)), reg, sel["n"], te, _ptr(t["tptr"]), _ptr(t["numel"]),
                                     _ptr(sel["param"]), _ptr(sel["local"]), _ptr(t["pm_off"]), _ptr(t["has_mask"]),
                                     _ptr(t["packed"]), _ptr(o["info"]), _ptr(o["W"]), _ptr(sel["proj"]), st)
                        _native.call("svdq_param_requantize", N, len(g.names), self.bits, self.stages,
                                     _ptr(sel["begin"]), _ptr(sel["proj"]), _ptr(t["present"]), _ptr(o["info"]),
                                     _ptr(o["coef"]), _ptr(o["chigh"]), _ptr(o["codes"]), _ptr(o["qscale"]),
                                     _ptr(o["qzp"]), _ptr(o["qres"]), _ptr(o["chat"]), st)

            def average(w_dev, order_dev):
                for g in self.groups.values():
                    for _, _, _, _, o in regions(g):
                        cl_dev, om_dev = self._cluster_tables
                        _native.call("svdq_param_average", N, len(g.names), _ptr(g.t["present"]), _ptr(w_dev),
                                     _ptr(order_dev), _ptr(cl_dev), _ptr(om_dev), _ptr(o["info"]), _ptr(o["chat"]),
                                     _ptr(o["W"]), _ptr(o["cbar"]), _ptr(o["gvec"]), _ptr(o["scal"]), st)

            any_exact = any(g.sel is not None for g in self.groups.values())
            if self.cluster_mode:
                # weights depend on the whole-model Gram: start its D2H on a side stream, run the
                # weight-independent solve meanwhile, then k-means on the host and the tiny average kernel
                self._cluster_begin()
                solve(None, self._order_dev)
                if any_exact:
                    project_exact()
                self._cluster_end()
                w_dev, order_dev = self._weights_table()
                average(w_dev, order_dev)
            elif any_exact:
                w_dev, order_dev = self._weights_table()
                solve(None, order_dev)
                project_exact()
                average(w_dev, order_dev)
            else:
                w_dev, order_dev = self._weights_table()
                solve(w_dev, order_dev)
            self._w_keep = (w_dev, order_dev)
            mark("k2")
            fused = self._fused_basis_buffers()             # None unless artifacts of <= 8 tasks are wanted
            for g in self.groups.values():
                t, tn = g.t, (g.tn or {})
                if fused is not None:
                    # the reference's default settings (diagnostics AND stored artifacts): pass 2 writes the bases too
                    fb = fused[g.dtype]
                    _native.call("svdq_basis_offsets", len(g.names), 0, te, _ptr(t["count"]), _ptr(t["tile_begin"]),
                                 _ptr(t["numel"]), _ptr(fb["row_off"]), st)
                    _native.call("svdq_reconstruct_merge_basis", _FLOAT_DTYPES[g.dtype], N, int(bool(cfg.svd_fp16)),
                                 int(bool(cfg.svd_center)), g.n_tiles, te, _ptr(t["tptr"]), _ptr(t["numel"]),
                                 _ptr(t["tile_param"]), _ptr(t["tile_local"]), _ptr(t["pm_off"]), _ptr(t["has_mask"]),
                                 _ptr(t["packed"]), _ptr(t["info"]), _ptr(t["W"]), _ptr(t["cbar"]), _ptr(t["gvec"]),
                                 _ptr(t["scal"]), _ptr(t["chat"]) if self.want_diag else None, _ptr(t["optr"]),
                                 _ptr(t["diag"]) if self.want_diag else None, _ptr(fb["row_off"]),
                                 _ptr(fb["uh_d"]), _ptr(fb["ul_d"]), _ptr(fb["mn_d"]), st)
                    continue
                _native.call("svdq_reconstruct_merge", _FLOAT_DTYPES[g.dtype], N, int(bool(cfg.svd_fp16)),
                             int(self.want_diag), int(bool(cfg.svd_center)), g.n_tiles, te, _ptr(t["tptr"]),
                             _ptr(t["numel"]), _ptr(t["tile_param"]), _ptr(t["tile_local"]), _ptr(t["pm_off"]),
                             _ptr(t["has_mask"]), _ptr(t["packed"]), _ptr(t["info"]), _ptr(t["W"]), _ptr(t["cbar"]),
                             _ptr(t["gvec"]), _ptr(t["scal"]), _ptr(t["chat"]), _ptr(t["optr"]), _ptr(t["diag"]),
                             _ptr(tn.get("info")), _ptr(tn.get("W")), _ptr(tn.get("cbar")), _ptr(tn.get("gvec")),
                             _ptr(tn.get("scal")), self.noise_shrink, st)
            mark("k3")
            if self.want_diag:
                for g in self.groups.values():
                    t = g.t
                    _native.call("svdq_diag_finalize", N, len(g.names), _ptr(t["diag"]), _ptr(t["tile_begin"]),
                                 _ptr(t["dm"]), _ptr(t["info"]), _ptr(t["gram_masked"]), _ptr(t["diag_out"]), st)
            mark("end")
        self._events = ev
        self._ran = True
        self._fetched = None
        self._bases_done = False
        return self

    @property
    def gpu_launches(self) -> int:
        """Kernels of libsvdq.so launched by one run()."""
        ng = len(self.groups)
        any_exact = any(g.sel is not None for g in self.groups.values())
        nreg = 2 if self.noise else 1
        # every svdq_gram_reduce is two launches (k2_gram_partial + k2_gram_reduce)
        n_reduce = 1 + (1 if self.wide and (self.cluster_mode or self.noise) else 0)
        return ng * (3 + n_reduce + nreg + (1 if self.want_diag else 0) + (nreg if (self.cluster_mode or any_exact) else 0)
                     + (2 * nreg if any_exact else 0))

    def event_times_ms(self) -> Dict[str, float]:
        ev = self._events
        torch.cuda.synchronize(self.device)
        names = ["start", "k1", "k2", "k3", "end"]
        return {f"{b}": ev[a].elapsed_time(ev[b]) for a, b in zip(names[:-1], names[1:])}

    # ------------------------------------------------------------------------------------------
    def _fetch(self) -> Dict[torch.dtype, Dict[str, np.ndarray]]:
        if not self._ran:
            self.run()
        if self._fetched is None:
            out = {}
            keys = ["info", "sv", "scal", "coef", "chigh", "codes", "qscale", "qzp", "qres", "cbar", "V", "dm"]
            for dt, g in self.groups.items():
                out[dt] = {k: g.t[k].cpu().numpy() for k in keys + (["diag_out"] if self.want_diag else [])}
                if g.tn is not None:      # noise region: same tables, key suffix "_n"
                    out[dt].update({k + "_n": g.tn[k].cpu().numpy() for k in keys})
            self._fetched = out
        return self._fetched

    def merged_state_dict(self, to_host: bool = False) -> "OrderedDict[str, torch.Tensor]":
        """merged = base + delta for every base key, in base key order (merge.py:483-494)."""
        if not self._ran:
            self.run()
        fetched = self._fetch()
        views: Dict[str, torch.Tensor] = {}
        for dt, g in self.groups.items():
            flat = g.t["out"]
            if to_host:
                if to_host == "reuse":      # one cached pinned buffer per size: results alias it until the next call
                    key = (flat.numel(), flat.dtype)
                    host = _PINNED.get(key)
                    if host is None:
                        host = _PINNED[key] = torch.empty(flat.shape, dtype=flat.dtype, pin_memory=True)
                else:
                    host = torch.empty(flat.shape, dtype=flat.dtype, pin_memory=True)
                host.copy_(flat, non_blocking=True)
                flat = host
            info = fetched[dt]["info"]
            for p, name in enumerate(g.names):
                if info[p, 0] != 0:
                    continue                      # no basis: the reference clones the base parameter
                views[name] = flat[g.out_off[p]: g.out_off[p] + g.numel[p]].view(g.shapes[p])
        if to_host:
            torch.cuda.synchronize(self.device)
        out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        for k in self.base_keys:
            if k in self.filtered_out:
                continue
            if k in views:
                out[k] = views[k]
            else:
                b = self.base_ref[k]
                out[k] = b.clone() if torch.is_tensor(b) else b
                if torch.is_tensor(b) and not to_host and b.device != self.device:
                    out[k] = out[k].to(self.device)
        return out

    # ------------------------------------------------------------------------------------------
    def _fused_basis_buffers(self):
        """Worst-case sized U_high / U_low / mean buffers for the fused write-out of pass 2 (svdq_reconstruct_merge_basis):
        k, r and the masked row count of a parameter are only known on the device when pass 2 is launched, so every
        parameter gets room for numel x N entries per block (views of the real size are cut after the fetch).
        None when the fused path does not apply (then K5 materialises the bases in a pass of its own)."""
        if not (self.materialize and not self.wide and self.N <= 8 and not self.noise
                and os.environ.get("SVDQ_FUSED_BASIS", "1") != "0"):
            return None
        if getattr(self, "_fused", None) is not None:
            return self._fused
        cfg, N = self.cfg, self.N
        udt = torch.float16 if cfg.svd_fp16 else torch.float32
        esz = 2 if cfg.svd_fp16 else 4
        need = sum(2 * int(sum(g.numel)) * N * esz + 4 * int(sum(g.numel)) for g in self.groups.values())
        free, _ = torch.cuda.mem_get_info(self.device)
        if need > free // 3:
            self._fused = None
            self.materialize_fused_skipped = f"worst-case basis buffers ({need / 1e9:.1f} GB) do not fit"
            return None
        out = {}
        for dt, g in self.groups.items():
            numel = np.asarray(g.numel, np.int64)
            al = 16 // esz
            nh = (numel * N + al - 1) // al * al
            nm = (numel + 3) // 4 * 4
            oh = np.concatenate([[0], np.cumsum(nh)]).astype(np.int64)
            om = np.concatenate([[0], np.cumsum(nm)]).astype(np.int64)
            uh = torch.empty(max(int(oh[-1]), 1), dtype=udt, device=self.device)
            ul = torch.empty(max(int(oh[-1]), 1), dtype=udt, device=self.device)
            mn = torch.empty(max(int(om[-1]), 1), dtype=torch.float32, device=self.device) if cfg.svd_center else None
            out[dt] = dict(uh=uh, ul=ul, mn=mn, oh=oh, ol=oh, om=om,
                           uh_d=_dev((uh.data_ptr() + oh[:-1] * esz).astype(np.int64), self.device),
                           ul_d=_dev((ul.data_ptr() + oh[:-1] * esz).astype(np.int64), self.device),
                           mn_d=_dev((mn.data_ptr() + om[:-1] * 4).astype(np.int64), self.device) if mn is not None else None,
                           row_off=torch.empty(max(g.n_tiles, 1), dtype=torch.int64, device=self.device))
        self._fused = out
        return out

    def _materialize_bases(self):
        """K5: U_high / U_low / mean compacted to the masked rows, in the artifact dtype (already written by pass 2
        when the fused write-out applied)."""
        if self._bases_done:
            return
        fetched = self._fetch()
        if getattr(self, "_fused", None) is not None:
            self._basis_store = {}
            for dt, fb in self._fused.items():
                info, dm = fetched[dt]["info"], fetched[dt]["dm"].astype(np.int64)
                self._basis_store[(dt, "masked")] = dict(
                    uh=fb["uh"], ul=fb["ul"], mn=fb["mn"], oh=fb["oh"], ol=fb["ol"], om=fb["om"], dm=dm,
                    r=info[:, 2].astype(np.int64), k=info[:, 3].astype(np.int64), ok=info[:, 0] == 0)
            self._bases_done = True
            return
        cfg, N, te = self.cfg, self.N, self.tile_elems
        st = _native.stream_ptr()
        udt = torch.float16 if cfg.svd_fp16 else torch.float32
        esz = 2 if cfg.svd_fp16 else 4
        # (dtype, region) -> flat U_high / U_low / mean buffers + per-parameter element offsets; the per-parameter
        # tensors handed out by basis_tensors() are views into them (one allocation instead of 3 P small ones)
        self._basis_store: Dict[Tuple[torch.dtype, str], Dict] = {}
        self._basis_aux: List[tuple] = []
        with torch.cuda.device(self.device):
            for dt, g in self.groups.items():
                t = g.t
                P = len(g.names)
                for reg in ((0, 1) if g.tn is not None else (0,)):
                    sfx, o = ("", g.t) if reg == 0 else ("_n", g.tn)
                    info, dm = fetched[dt]["info" + sfx], fetched[dt]["dm" + sfx].astype(np.int64)
                    ok = info[:, 0] == 0
                    r, k = info[:, 2].astype(np.int64), info[:, 3].astype(np.int64)
                    al = 16 // esz                                    # slices start 16-byte aligned
                    nh = np.where(ok, (dm * k + al - 1) // al * al, 0)
                    nl = np.where(ok, (dm * (r - k) + al - 1) // al * al, 0)
                    nm = np.where(ok, (dm + 3) // 4 * 4, 0)
                    oh, ol, om = (np.concatenate([[0], np.cumsum(x)]).astype(np.int64) for x in (nh, nl, nm))
                    uh = torch.empty(max(int(oh[-1]), 1), dtype=udt, device=self.device)
                    ul = torch.empty(max(int(ol[-1]), 1), dtype=udt, device=self.device)
                    mn = torch.empty(max(int(om[-1]), 1), dtype=torch.float32, device=self.device) \
                        if cfg.svd_center else None
                    uh_d = _dev(np.where(ok, uh.data_ptr() + oh[:-1] * esz, 0).astype(np.int64), self.device)
                    ul_d = _dev(np.where(ok, ul.data_ptr() + ol[:-1] * esz, 0).astype(np.int64), self.device)
                    mn_d = _dev(np.where(ok, mn.data_ptr() + om[:-1] * 4, 0).astype(np.int64), self.device) \
                        if mn is not None else None
                    row_off = torch.empty(max(g.n_tiles, 1), dtype=torch.int64, device=self.device)
                    _native.call("svdq_basis_offsets", P, reg, te, _ptr(t["count"]), _ptr(t["tile_begin"]),
                                 _ptr(t["numel"]), _ptr(row_off), st)
                    _native.call("svdq_write_basis", _FLOAT_DTYPES[g.dtype], N, int(bool(cfg.svd_fp16)),
                                 int(bool(cfg.svd_center)), reg, g.n_tiles, te, _ptr(t["tptr"]), _ptr(t["numel"]),
                                 _ptr(t["tile_param"]), _ptr(t["tile_local"]), _ptr(t["pm_off"]), _ptr(t["has_mask"]),
                                 _ptr(t["packed"]), _ptr(o["info"]), _ptr(o["W"]), _ptr(row_off), _ptr(uh_d),
                                 _ptr(ul_d), _ptr(mn_d), st)
                    self._basis_aux.append((row_off, uh_d, ul_d, mn_d))      # alive until the next materialisation
                    self._basis_store[(dt, "masked" if reg == 0 else "noise")] = dict(
                        uh=uh, ul=ul, mn=mn, oh=oh, ol=ol, om=om, dm=dm, r=r, k=k, ok=ok)
        self._bases_done = True

    def basis_tensors(self, dt: torch.dtype, p: int, region: str = "masked"):
        """(U_high [D x k], U_low [D x (r-k)], mean [D x 1] | None) of parameter p of dtype group dt."""
        self._materialize_bases()
        st = self._basis_store[(dt, region)]
        if not st["ok"][p]:
            return None
        d, r, k = int(st["dm"][p]), int(st["r"][p]), int(st["k"][p])
        oh, ol, om = int(st["oh"][p]), int(st["ol"][p]), int(st["om"][p])
        uh = st["uh"][oh: oh + d * k].view(d, k)
        ul = st["ul"][ol: ol + d * (r - k)].view(d, r - k)
        mn = st["mn"][om: om + d].view(d, 1) if st["mn"] is not None else None
        return uh, ul, mn

    def combined_masks(self) -> Dict[str, torch.Tensor]:
        """Combined tall masks as torch.bool tensors (what combine_masks returns)."""
        if not self._ran:
            self.run()
        st = _native.stream_ptr()
        out = {}
        with torch.cuda.device(self.device):
            for g in self.groups.values():
                hm = g.host["has_mask"]
                pm_off = g.t["pm_off"].cpu().numpy()
                for p, name in enumerate(g.names):
                    if not hm[p]:
                        continue
                    m = torch.empty(g.numel[p], dtype=torch.bool, device=self.device)
                    _native.call("svdq_unpack_mask", g.t["packed"].data_ptr() + 4 * int(pm_off[p]), g.numel[p],
                                 m.data_ptr(), st)
                    out[name] = m.view(g.shapes[p])
        return out

    # ------------------------------------------------------------------------------------------
    def results(self, to_host: bool = False) -> Dict:
        """{"merged_state_dict", "diagnostics", "bases", "compressed"} as cli.py:773-778."""
        from .results import LazyBases, LazyCompressed, build_diagnostics
        merged = self.merged_state_dict(to_host=to_host)
        fetched = self._fetch()
        if self.materialize:
            self._materialize_bases()
        index = {}
        for dt, g in self.groups.items():
            info = fetched[dt]["info"]
            for p, name in enumerate(g.names):
                if info[p, 0] == 0:
                    index[name] = (dt, p)
        bases = LazyBases(self, index)
        compressed = LazyCompressed(self, index)
        if self.want_diag:
            diag = build_diagnostics(self, index)
        else:
            diag = {}
        diag["task_weights"] = self.weights
        if self.cluster_assignments:
            diag["cluster_assignments"] = self.cluster_assignments
        return {"merged_state_dict": merged, "diagnostics": diag, "bases": bases, "compressed": compressed}


def merge_state_dicts(base, finetuned, task_masks, config, device: Optional[str] = None, **kw) -> Dict:
    """Fused fast path: the whole of run_svd_hybrid_pipeline steps 1-9 (cli.py:146-722) on in-memory
    state dicts.  Returns the reference's result dict (cli.py:773-778)."""
    import time as _time
    to_host = kw.pop("to_host", False)
    job = MergeJob(base, finetuned, task_masks, config, device, **kw)
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
    res["job"] = job
    return res


def task_vector_gram(task_vectors: Mapping[str, Mapping[str, torch.Tensor]], names: Sequence[str]) -> np.ndarray:
    """Whole-model N x N Gram of already-formed task vectors via K1 (base = zeros)."""
    _native.require_cuda()
    dev = torch.device("cuda")
    N = len(names)
    if N > MAX_STREAM_TASKS:
        raise ValueError(f"task_vector_gram: at most {MAX_STREAM_TASKS} task vectors")
    params = sorted({p for n in names for p in task_vectors[n].keys()})
    total = np.zeros((N, N), np.float64)
    from .svd_hybrid.config import SVDHybridConfig
    cfg = SVDHybridConfig(tasks=list(names), svd_weighting="cluster", svd_eval_reconstruction=False)
    ref = {}
    for p in params:
        src = next(task_vectors[n][p] for n in names if p in task_vectors[n])
        ref[p] = torch.zeros_like(src, device=dev)
    fts = {n: {p: v.to(dev) for p, v in task_vectors[n].items()} for n in names}
    job = MergeJob(ref, fts, None, cfg, "cuda", cluster_assignments={n: 0 for n in names})
    st = _native.stream_ptr()
    for g in job.groups.values():
        t = g.t
        _native.call("svdq_tv_mask_gram", _FLOAT_DTYPES[g.dtype], N, 0, 1, g.n_tiles, job.tile_elems,
                     _ptr(t["tptr"]), None, _ptr(t["numel"]), _ptr(t["tile_param"]), _ptr(t["tile_local"]),
                     _ptr(t["pm_off"]), _ptr(t["packed"]), _ptr(t["gram"]), _ptr(t["count"]), st)
        _native.call("svdq_gram_reduce", N, 1, len(g.names), 0, _ptr(t["gram"]), _ptr(t["count"]),
                     _ptr(t["tile_begin"]), None, None, _ptr(t["gram_masked"]), _ptr(t["gram_all"]), _ptr(t["dm"]),
                     None, None, st)
        total += g.t["gram_all"].sum(dim=0).view(N, N).cpu().numpy()
    return total


def factor_columns(vectors: Sequence[torch.Tensor], center: bool = True, energy_threshold: float = 0.90,
                   max_rank: Optional[int] = None, fp16: bool = False, sign_ref=None) -> Dict:
    """Thin SVD + rank selection of the [D x N] matrix whose columns are ``vectors`` (K1 with a zero
    base -> K2 -> K5).  Backs construct_basis / compute_svd of the fine-grained API.

    -> {"U_high","U_low","singular_values","k","mean","energy_retained","D","N","V","coef"}; V[t][j]
    are the right singular vectors (fp64), coef[t][j] the closed-form task coefficients."""
    from .svd_hybrid.config import SVDHybridConfig
    _native.require_cuda()
    if len(vectors) == 0:
        raise ValueError("Empty delta list")
    dev = torch.device("cuda")
    cols = [v.detach().to(dev, torch.float32).contiguous().view(-1) for v in vectors]
    D = cols[0].numel()
    names = [f"c{i:02d}" for i in range(len(cols))]
    cfg = SVDHybridConfig(tasks=names, svd_center=bool(center), svd_energy_threshold=float(energy_threshold),
                          svd_max_rank=max_rank, svd_fp16=bool(fp16), svd_store_artifacts=False,
                          svd_eval_reconstruction=False)
    base = {"x": torch.zeros(D, dtype=torch.float32, device=dev)}
    fts = {n: {"x": c} for n, c in zip(names, cols)}
    job = MergeJob(base, fts, None, cfg, "cuda", diagnostics=False, materialize_bases=True,
                   sign_ref={"x": sign_ref} if sign_ref is not None else None)
    job.run()
    res = job.results()
    if "x" not in res["bases"]:
        raise ValueError("Empty delta list")
    b = dict(res["bases"]["x"]["masked"])
    b["V"] = torch.from_numpy(res["bases"].right_vectors("x").copy())
    b["coef"] = torch.from_numpy(res["compressed"].raw_coefficients("x").copy())
    return b
