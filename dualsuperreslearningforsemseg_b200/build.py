"""Builds libdsrl_b200.so in-tree with nvcc for sm_100a (the only supported target).

The .so is git-ignored but travels to the GPU box with the repo snapshot.  `python -m
dualsuperreslearningforsemseg_b200.build` or `__graft_entry__.build()` runs this.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libdsrl_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

OBJ_DIR = os.path.join(PKG_DIR, "_obj")          # git-ignored; objects are rebuilt on the GPU box only if sources changed

COMPILE_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]
NVCC_FLAGS = COMPILE_FLAGS + ["-shared"]          # the single-command equivalent (profiles/, docs)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _obj_path(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")


def build(force: bool = False, verbose: bool = False) -> str:
    """One nvcc -c per translation unit, in parallel (only the ones whose sources / headers changed), then one link."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libdsrl_b200.so must be built where the CUDA toolkit is installed")
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)
    procs = []
    for src in sources():
        obj = _obj_path(src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            continue
        cmd = [nvcc, *COMPILE_FLAGS, "-I", INCLUDE, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    failed = []
    for src, p in procs:
        out, err = p.communicate()
        if p.returncode != 0:
            failed.append(f"{src}:\n{out}{err}")
        elif verbose:
            print(err, file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed))
    cmd = [nvcc, *LINK_FLAGS, "-o", LIB_PATH + ".tmp", *[_obj_path(s) for s in sources()]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
