"""In-tree build of the CUDA library (sm_100a only).

`python -m arxiv_rag_b200.build` (or `__graft_entry__.build()`) compiles every `csrc/*.cu`
with nvcc for `compute_100a/sm_100a` and links `lib/libarxiv_rag_b200.so`. nvcc cross-compiles
without a GPU, so this runs on the CPU-only dev box; the .so travels to the GPU box in-tree.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "build"
LIB = LIBDIR / "libarxiv_rag_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None,
          debug: bool = False, variant: str | None = None) -> Path:
    """debug=True builds lib/libarxiv_rag_b200_dbg.so with -DARB_HANG_GUARD: a stuck mbarrier
    wait traps (with a printf) instead of spinning forever — used for first runs of new kernels.
    variant='name' builds lib/libarxiv_rag_b200_<name>.so with `extra_flags` (A/B builds selected at
    run time through ARB_LIB_PATH)."""
    global OBJDIR, LIB
    if variant:
        OBJDIR = PKG / "build" / f"variant_{variant}"
        OBJDIR.mkdir(parents=True, exist_ok=True)
        LIB = LIBDIR / f"libarxiv_rag_b200_{variant}.so"
    elif debug:
        OBJDIR = PKG / "build_dbg"
        LIB = LIBDIR / "libarxiv_rag_b200_dbg.so"
        extra_flags = (extra_flags or []) + ["-DARB_HANG_GUARD"]
    else:
        OBJDIR = PKG / "build"
        LIB = LIBDIR / "libarxiv_rag_b200.so"
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + \
        [PKG.parent / "include" / "arxiv_rag_b200.h"]
    sources = sorted(CSRC.glob("*.cu"))
    flags = CFLAGS + (extra_flags or [])

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC, *ARCH, *flags, "-c", str(src), "-o", str(obj)]
            res = subprocess.run(cmd, capture_output=True, text=True)
            (OBJDIR / (src.stem + ".log")).write_text(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
            if verbose:
                print(res.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or _stale(LIB, objs):
        # shared cudart: the process already has one (torch's), and a static copy would carry the
        # runtime's whole entry-point table into the shipped artifact
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "shared"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    variant = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), None)
    flags = [a for a in sys.argv[1:] if a.startswith("-D")]
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv,
                 variant=variant, extra_flags=flags or None)
    print(path)
