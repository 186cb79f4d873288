"""In-tree build of libsvdq.so (sm_100a) with nvcc; no torch headers, no JIT cache.

``python -m svd_quantization_task_merging_b200.build`` or ``__graft_entry__.build()``.
The library travels to the GPU box with the repository snapshot (``*.so`` is git-ignored, not
gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libsvdq.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -Xfatbin -compress-all: the library instantiates every (kernel family x task count x dtype x flag) combination; the
# compressed device images are a third of the size (the library travels to the GPU box with every snapshot)
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xfatbin", "-compress-all"] + ARCH_FLAGS

# (source, SVDQ_DTYPE or None)
UNITS: List[Tuple[str, object]] = (
    [("k1_tv_mask_gram.cu", d) for d in (0, 1, 2)]
    + [("k3_reconstruct_merge.cu", d) for d in (0, 1, 2)]
    + [("k3c_merge_diag_compact.cu", d) for d in (0, 1, 2)]
    + [("k5_basis_misc.cu", d) for d in (0, 1, 2)]
    + [("k1s_tv_mask_gram_staged.cu", d) for d in (0, 1, 2)]
    + [("k3s_reconstruct_merge_staged.cu", d) for d in (0, 1, 2)]
    + [("k6_wide.cu", d) for d in (0, 1, 2)]
    + [("k7_project_exact.cu", d) for d in (0, 1, 2)]
    + [("k8_gram_staged.cu", d) for d in (0, 1, 2)]
    + [("k9_gram_tc.cu", d) for d in (0, 1, 2)]
    + [("k10_merge_tc.cu", d) for d in (0, 1, 2)]
    + [("k12_gram_wide_tc.cu", 0), ("k13_merge_wide_tc.cu", 0)]
    + [("k2_param_solve.cu", None), ("k11_reload_merge.cu", None), ("k14_operators.cu", None), ("k4_rtvq_large.cu", None), ("svdq_capi.cu", None), ("host_kmeans.cpp", None), ("host_pack.cpp", None)]
)
HEADERS = ["svdq_common.cuh", "svdq_kernels.h", "k2_core.h", "k3_body.cuh", "stage_pipe.cuh", os.path.join("..", "..", "include", "svdq.h")]


def _digest(paths: List[str], extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _includes(path: str, seen=None) -> List[str]:
    """The file and every in-tree header it includes with quotes, transitively (the unit's real dependencies: a
    change to one header rebuilds only the units that see it)."""
    import re
    seen = seen if seen is not None else []
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.append(path)
    with open(path) as f:
        for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', f.read(), flags=re.M):
            _includes(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _compile(unit: Tuple[str, object], verbose: bool) -> str:
    src, dt = unit
    stem = os.path.splitext(src)[0] + ("" if dt is None else f"_dt{dt}")
    obj = os.path.join(OBJDIR, stem + ".o")
    stamp = obj + ".sha"
    deps = sorted(_includes(os.path.join(CSRC, src)))
    flags = COMMON + ([] if dt is None else [f"-DSVDQ_DTYPE={dt}"])
    if src.endswith(".cpp"):                       # host-only unit: keep float32 arithmetic un-fused
        flags = flags + ["-Xcompiler", "-ffp-contract=off"]
    want = _digest(deps, " ".join(flags))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj
    cmd = [NVCC] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src} (dtype {dt}):\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want)
    return obj


def build(force: bool = False, verbose: bool = False, jobs: int = 0) -> str:
    """Compile every CUDA translation unit for sm_100a and link libsvdq.so.  Returns its path."""
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    jobs = jobs or min(len(UNITS), os.cpu_count() or 4)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        objs = list(ex.map(lambda u: _compile(u, verbose), UNITS))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ARCH_FLAGS + ["-Xcompiler", "-fPIC", "-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
