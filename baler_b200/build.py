"""Build libbaler_b200.so in-tree with nvcc for sm_100a (the built .so travels to the GPU box).

    python -m baler_b200.build [--force] [--verbose]

Every translation unit is compiled to its own object (in parallel, only when stale) and linked into the shared library.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libbaler_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in sources() + headers())


def _compile(src, verbose):
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", _obj(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max([os.path.getmtime(p) for p in headers()] + [0.0])
    todo = [s for s in sources()
            if force or not os.path.exists(_obj(s)) or os.path.getmtime(_obj(s)) < max(os.path.getmtime(s), hdr_t)]
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as pool:
        for src, rc, log in pool.map(lambda s: _compile(s, verbose), todo):
            if verbose or rc != 0:
                sys.stderr.write(log)
            if rc != 0:
                raise RuntimeError("nvcc failed on %s" % src)
    r = subprocess.run([NVCC, "-shared", "-Xcompiler", "-pthread"] + [_obj(s) for s in sources()] + ["-o", LIB],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("linking libbaler_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
