"""Build libphet_b200.so in-tree with nvcc for sm_100a (no torch involvement).

    python -m phet_b200.build [--force] [--verbose]

The shared library lands next to this file (phet_b200/libphet_b200.so); it is git-ignored but
travels to the GPU box with the gpurun snapshot.  Cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libphet_b200.so")
STAMP = LIB + ".stamp"
SOURCES = ["api.cu", "normalize.cu", "finalize.cu", "classstats.cu", "rank.cu", "mtx.cu", "anova.cu", "step3big.cu", "step1_f32.cu", "step1_f64.cu", "step3_f32.cu", "step3_f64.cu",
           "step3p_f32.cu", "step3p_f64.cu", "step3s_f32.cu", "step3s_f64.cu", "fy_apply.cu", "replay.cu", "mt_replay.cpp"]
# host-only sources go through g++ directly (AVX-512 intrinsics behind target attributes + a runtime CPU check)
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-Wall", "-fno-math-errno"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", 
              "--expt-relaxed-constexpr"] + os.environ.get("PHET_NVCC_EXTRA", "").split()


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _source_hash():
    """Content hash of every input of the build (mtimes do not survive the gpurun snapshot)."""
    import hashlib
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS + SOURCES).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _source_hash()


def _deps(path, seen=None):
    """Transitive closure of the quoted #includes of a source (paths relative to the including file)."""
    import re
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as fh:
        for inc in re.findall(r'^\s*#include\s+"([^"]+)"', fh.read(), flags=re.M):
            _deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _object_hash(src):
    import hashlib
    h = hashlib.sha256(" ".join(CXX_FLAGS if src.endswith(".cpp") else NVCC_FLAGS).encode())
    for f in sorted(_deps(os.path.join(CSRC, src))):
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".stamp"
    want = _object_hash(src)
    if not verbose and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return obj, ""   # this object's source and headers are unchanged
    if src.endswith(".cpp"):
        cmd = [shutil.which("g++") or "g++", *CXX_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    else:
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose and not src.endswith(".cpp"):
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compile failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(want)
    return obj, r.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if verbose:
        with open(os.path.join(BUILD, "ptxas.log"), "w") as fh:
            for _, err in results:
                fh.write(err)
    cmd = [_nvcc(), "-shared", "-o", LIB, *[o for o, _ in results], "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(STAMP, "w") as fh:
        fh.write(_source_hash())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
