"""In-tree build of libvpf_b200.so (nvcc, sm_100a only, -lineinfo).

    python -m vipformer_b200.build [--force] [--verbose]

The .so lands next to this file so that it travels to the GPU box with the
repo snapshot.  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libvpf_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("VPF_NVCC_EXTRA", "").split()   # e.g. -DVPF_GEMM_TIMING for tools/gemm_phase_timing.py


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + ["../../include/vpf.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and (f.endswith((".cuh", ".h")) or os.path.abspath(p) == os.path.abspath(path)):
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src, force, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp = obj + ".sha1"
    dig = _digest(path)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False, ""
    cmd = ["nvcc", *NVCC_FLAGS, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(dig)
    return obj, True, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), sources()))
    objs = [o for o, _, _ in res]
    if verbose:
        for _, _, log in res:
            if log:
                sys.stderr.write(log)
    if force or any(ch for _, ch, _ in res) or not os.path.exists(SO):
        cmd = ["nvcc", "-shared", "-o", SO, *objs, "-lcuda", "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
