"""In-tree nvcc build of libsceneego_b200.so (sm_100a only, no torch headers).

    python -m sceneego_b200.build [--force]

The library exposes only the C-ABI of include/sceneego_b200.h; Python binds it
with ctypes (sceneego_b200/_lib.py).  nvcc cross-compiles without a GPU.

Every source is compiled to its own object (in parallel) and linked; a SHA-256 over the sources, the header and
the flags is stored next to the library (`libsceneego_b200.so.srchash`), and the library is rebuilt whenever that
hash differs from the current tree's -- a stale prebuilt `.so` shipped with a snapshot is never silently reused.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsceneego_b200.so")            # bf16 activations / weights (default)
LIB_F16 = os.path.join(HERE, "libsceneego_b200_f16.so")    # the same sources with -DSCENEEGO_ACT_F16 (fp16 storage)
HASH_FILE = LIB + ".srchash"
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["geometry.cu", "softargmax.cu", "v2v.cu", "stem.cu", "stem_march.cu", "tail.cu", "march.cu", "eval.cu", "handoff.cu", "feature_conv.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def source_hash() -> str:
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    deps.append(os.path.join(HERE, "..", "include", "sceneego_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    h.update(" ".join(NVCC_FLAGS + _sources()).encode())
    return h.hexdigest()


def recorded_hash() -> str:
    try:
        return open(HASH_FILE).read().strip()
    except OSError:
        return ""


def _stale() -> bool:
    return not (os.path.exists(LIB) and os.path.exists(LIB_F16)) or recorded_hash() != source_hash()


VARIANTS = [("bf16", LIB, []), ("f16", LIB_F16, ["-DSCENEEGO_ACT_F16"])]


def build(force: bool = False, verbose: bool = False) -> str:
    """Build both libraries (bf16 and fp16 activation storage) from the same sources; returns the default one."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    want = source_hash()

    def compile_one(job):
        tag, src, extra = job
        obj = os.path.join(OBJ_DIR, src.replace(".cu", f".{tag}.o"))
        r = subprocess.run([nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return tag, src, obj, r

    jobs = [(tag, src, extra) for tag, _, extra in VARIANTS for src in _sources()]
    jobs.sort(key=lambda j: -os.path.getsize(os.path.join(CSRC, j[1])))          # the long compiles first
    objs = {tag: [] for tag, _, _ in VARIANTS}
    failed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for tag, src, obj, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"---- {src} [{tag}]\n{r.stdout}{r.stderr}")
            failed |= r.returncode != 0
            objs[tag].append(obj)
    if failed:
        raise RuntimeError("nvcc failed building libsceneego_b200")
    for tag, lib, _ in VARIANTS:
        r = subprocess.run([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs[tag],
                           capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed linking {os.path.basename(lib)}")
    with open(HASH_FILE, "w") as f:
        f.write(want + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-q" not in sys.argv))
