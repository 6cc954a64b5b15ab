"""Build libgu_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m griduniverse_b200.build [--force]

The library is written next to the package (griduniverse_b200/lib/libgu_b200.so) so it
travels to the GPU box with the repo snapshot; *.so is git-ignored.
"""
import glob
import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libgu_b200.so")
STAMP = os.path.join(LIB_DIR, "libgu_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                       # the reference's arithmetic has no fused multiply-add
    "--shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    """Content hash of the sources (paths relative to the repo root, so any checkout agrees)."""
    h = hashlib.sha256()
    for p in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h"))):
        h.update(os.path.relpath(p, ROOT).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(out_path, defines):
    """Developer helper: build a tuning variant (extra -D flags) to another path."""
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-D" + d for d in defines] + \
        ["-I", INCLUDE, "-I", CSRC, "-o", out_path] + _sources()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return out_path


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns the library path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    # one nvcc -c per translation unit, in parallel (no cross-file device calls), then one link
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "--shared"]
    jobs = []
    for src in _sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc()] + cflags + ["-I", INCLUDE, "-I", CSRC, "-c", "-o", obj, src]
        jobs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, failed = "", False
    for cmd, obj, proc in jobs:
        out = proc.communicate()[0]
        log += " ".join(cmd) + "\n" + out
        failed = failed or proc.returncode != 0
    if not failed:
        cmd = [_nvcc(), "--shared", "-o", LIB_PATH] + [obj for _, obj, _ in jobs]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log += " ".join(cmd) + "\n" + proc.stdout
        failed = proc.returncode != 0
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libgu_b200.so")
    if verbose:
        print(log)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose=True)
    print("built", path)
