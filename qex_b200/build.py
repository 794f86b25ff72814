"""Build recipe for libqexxc.so (hand-written CUDA for sm_100a) -- in-tree, no JIT cache.

``python -m qex_b200.build`` compiles every translation unit under ``qex_b200/csrc`` with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` and links ``qex_b200/libqexxc.so``.
nvcc cross-compiles without a GPU, so this also runs on the CPU-only builder box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
BUILD = HERE / "_build"
LIB = HERE / "libqexxc.so"
SOURCES = ["api.cu", "contract.cu", "contract_i8.cu", "ao.cu", "pointwise.cu", "xc_mlp.cu", "xc_mlp_tc.cu", "xc_mlp_wide.cu", "xc_global.cu", "xc_qnn.cu", "jk.cu", "eigh.cu", "grid.cu", "xc_lda.cu", "comm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> Path:
    BUILD.mkdir(exist_ok=True)
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "qexxc.h"]
    stamp = BUILD / "stamp"
    dig = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        if verbose:
            print(f"build: {LIB.name} is up to date (source digest {dig[:12]} matches {stamp}); "
                  f"`python -m qex_b200.build --force` recompiles")
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str):
        obj = BUILD / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (src + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr[-4000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    stamp.write_text(dig)
    # what was built, for the record (the .so itself is git-ignored and travels to the GPU box with the snapshot)
    import json
    import time
    ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    (BUILD / "build_info.json").write_text(json.dumps({
        "library": LIB.name, "sources": SOURCES, "nvcc": ver, "flags": NVCC_FLAGS, "source_digest": dig,
        "built_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "bytes": LIB.stat().st_size}, indent=1))
    if verbose:
        print(f"built {LIB} ({len(SOURCES)} translation units, {ver})")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
