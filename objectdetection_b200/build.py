"""Builds ``objectdetection_b200/libodhead.so`` in-tree with nvcc for sm_100a (B200) only.

    python -m objectdetection_b200.build [--force] [--verbose]

Flags that matter for parity: ``-fmad=false`` (no FMA contraction, so fp32 results follow the reference's
operation order), default IEEE ``-prec-div/-prec-sqrt``, no fast-math.  ``-lineinfo`` keeps ncu's source page
usable.  The object files go to ``csrc/_build`` (git-ignored); the ``.so`` is git-ignored too but travels with
``gpurun`` snapshots.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libodhead.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "-Xptxas", "-warn-spills",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libodhead.so for sm_100a)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def source_hash() -> str:
    """sha1 over the CUDA sources, their headers and the C ABI header (sorted by name). Compiled into the library
    (``od_source_hash``) and compared by ``_lib.lib()``, so a stale .so is refused instead of silently tested."""
    import hashlib
    h = hashlib.sha1()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(INCLUDE, "odhead.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def needs_build() -> bool:
    """True when libodhead.so is missing or was built from other sources than the tree holds (hash, not mtimes:
    file times do not survive a snapshot copy to another box)."""
    stamp = os.path.join(OBJ, "source_hash.txt")
    if not os.path.exists(LIB) or not os.path.exists(stamp):
        return True
    return open(stamp).read().strip() != source_hash()


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, defines) -> str:
    """Developer aid: a second library with extra -D flags (A/B runs select it with ODHEAD_LIB)."""
    nvcc = _nvcc()
    out = os.path.join(HERE, f"libodhead_{name}.so")
    tmp = os.path.join(OBJ, f"variant_{name}")
    os.makedirs(tmp, exist_ok=True)
    objs = []
    for src in sources():
        obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
        subprocess.run([nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], f'-DOD_SOURCE_HASH="{source_hash()}"', "-I", INCLUDE,
                        "-c", src, "-o", obj], check=True, capture_output=True, text=True)
        objs.append(obj)
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs, "-Xcompiler", "-fPIC",
                    "-Xlinker", "--no-undefined", "-lcudart_static", "-lpthread", "-ldl", "-lrt"], check=True,
                   capture_output=True, text=True)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "odhead.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    digest = source_hash()
    stamp = os.path.join(OBJ, "source_hash.txt")
    old = open(stamp).read().strip() if os.path.exists(stamp) else ""
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        is_core = os.path.basename(src) == "core.cu"
        if force or _stale(obj, [src] + headers) or (is_core and old != digest):
            cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
            if is_core:
                cmd.insert(1, f'-DOD_SOURCE_HASH="{digest}"')
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    failed = False
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, r in ex.map(run, jobs):
            out = (r.stdout + r.stderr).strip()
            if r.returncode != 0:
                failed = True
                print(f"[odhead build] FAILED {os.path.basename(src)}\n{out}", file=sys.stderr)
            elif out and (verbose or "warning" in out.lower()):
                print(f"[odhead build] {os.path.basename(src)}\n{out}", file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed")
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources()]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
               "-Xcompiler", "-fPIC", "-Xlinker", "--no-undefined", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr, file=sys.stderr)
            raise RuntimeError("link failed")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
