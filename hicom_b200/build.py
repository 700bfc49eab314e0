"""In-tree build of libhicom_b200.so (sm_100a only).

    python -m hicom_b200.build [--force]

nvcc cross-compiles without a GPU.  The library lands in ``hicom_b200/lib/`` (git-ignored, but it
travels with gpurun snapshots).  Objects are cached under ``build/`` keyed by a hash of each
source plus the headers.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhicom_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
SOURCES = ["api.cu", "local_attend.cu", "gemm_simt.cu", "rowwise.cu", "gemm_tc.cu", "skinny.cu", "backward.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _header_digest():
    h = hashlib.sha256()
    paths = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    paths.append(os.path.join(ROOT, "include", "hicom_b200.h"))
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src, hdr_digest, force):
    path = os.path.join(CSRC, src)
    with open(path, "rb") as f:
        digest = hashlib.sha256(f.read() + hdr_digest.encode()).hexdigest()[:16]
    obj = os.path.join(OBJ_DIR, f"{os.path.splitext(src)[0]}.{digest}.o")
    if force or not os.path.exists(obj):
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    hd = _header_digest()
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, hd, force), SOURCES))
    for f in os.listdir(OBJ_DIR):  # drop objects of older source versions
        if f.endswith(".o") and os.path.join(OBJ_DIR, f) not in objs:
            os.remove(os.path.join(OBJ_DIR, f))
    stamp = os.path.join(LIB_DIR, ".stamp")
    want = "\n".join(objs)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB_PATH
    tmp = LIB_PATH + f".tmp{os.getpid()}"  # link beside the target, then rename: readers never see a half-written library
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs,
           "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    with open(stamp, "w") as f:
        f.write(want)
    if verbose:
        print("built", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
