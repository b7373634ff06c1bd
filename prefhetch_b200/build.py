"""Build the sm_100a shared library (C ABI) in-tree with nvcc.  No JIT cache: the .so sits next to
the package so it travels to the GPU box with the snapshot."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libprefhetch_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-Xptxas", "-v", "--expt-relaxed-constexpr",
    "-ccbin", "/usr/bin/g++",
]


def sources():
    return sorted(CSRC.glob("*.cu")), sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                                             [PKG.parent / "include" / "prefhetch_b200.h"])


def needs_build() -> bool:
    if not LIB.exists():
        return True
    cu, hdr = sources()
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in cu + hdr)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cu, _ = sources()
    cmd = [NVCC, *FLAGS, "-o", str(LIB), *[str(c) for c in cu], "-lz", "-ldl"]  # zlib: SEAL compr_mode zlib streams; dl: libzstd.so.1 bound at run time
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = PKG / "build.log"
    log.write_text(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError(f"nvcc failed (see {log})")
    return LIB


def build_variant(name: str, defines) -> Path:
    """an experiment build next to the product library: libprefhetch_b200.<name>.so compiled with extra -D flags
    (select it with PF_LIB=<path>); used for kernel A/B measurements recorded in profiles/README.md"""
    out = PKG / f"libprefhetch_b200.{name}.so"
    cu, _ = sources()
    cmd = [NVCC, *FLAGS, *defines, "-o", str(out), *[str(c) for c in cu], "-lz", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (PKG / f"build.{name}.log").write_text(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for variant {name}")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        build(force="--force" in sys.argv, verbose=True)
        print(LIB)
