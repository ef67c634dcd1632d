"""Builds visapi_b200/librectpack_b200.so (sm_100a only) with nvcc, in tree.

    python -m visapi_b200.build [--force]

The shared library is git-ignored but travels to the GPU box with the snapshot.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(HERE, "librectpack_b200.so")
SOURCES = ["api_basic.cu", "flat_search.cu", "puct.cu", "resident_k22.cu", "resident_k24.cu", "resident_k42.cu",
           "resident_k44.cu"]
HEADERS = ["board.cuh", "philox.cuh", "common.cuh", "lanes.cuh", "search_types.cuh", "search_queue.cuh",
           "search_kernels.cuh", "search_resident.cuh", "resident_variant.cuh",
           os.path.join("..", "..", "include", "rectpack_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--threads", "2",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; rectpack_b200 cannot be built")
    return exe


def _mtime(path):
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


HASH_PATH = LIB_PATH + ".srchash"


def source_hash():
    """Digest of every source the library is built from (and of the flags): the library is
    rebuilt when this changes, not when file times do — a snapshot copied to the GPU box keeps
    its prebuilt .so even if the copy reset the mtimes."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in sorted(_sources()) + sorted(HEADERS):
        path = os.path.join(CSRC, name)
        if os.path.exists(path):
            with open(path, "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    return h.hexdigest()


def stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def _compile(src):
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    cmd = [nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, proc.stdout, proc.stderr))
    return obj


def build(force=False, verbose=True):
    if not force and not stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    if verbose:
        print("[visapi_b200] nvcc sm_100a: %s" % ", ".join(srcs), file=sys.stderr)
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        objs = list(pool.map(_compile, srcs))
    cmd = [nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                        "-cudart", "static"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (proc.stdout, proc.stderr))
    with open(HASH_PATH, "w") as f:
        f.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
