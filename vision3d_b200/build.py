"""Build vision3d_b200/libv3d_b200.so (sm_100a only) with nvcc, in-tree.

    python -m vision3d_b200.build [--force] [-v]

One translation unit per kernel family; the IoU/NMS unit is compiled with -fmad=false because its
results must be bit-identical to the reference arithmetic evaluated without fused multiply-add
(see csrc/iou_nms.cu). nvcc cross-compiles without a GPU; the .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libv3d_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

COMMON = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
          "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
UNITS = {
    "cabi.cu": [],
    "iou_nms.cu": ["-fmad=false"],
    "voxelize.cu": [],
    "rulebook.cu": [],
    "sparse_conv.cu": [],
    "sparse_conv_bwd.cu": [],
    "sparse_conv_tc.cu": [],
    "dense.cu": [],
    "head.cu": ["-fmad=false"],
    "pointops.cu": [],
    "sa_fused.cu": [],
}


def _sources():
    return [u for u in UNITS if os.path.exists(os.path.join(CSRC, u))]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "v3d_b200.h"))
    jobs = []
    for u in _sources():
        src = os.path.join(CSRC, u)
        obj = os.path.join(OBJ, u.replace(".cu", ".o"))
        if force or _stale(obj, [src] + headers):
            jobs.append((u, [NVCC] + COMMON + UNITS[u] + ["-c", src, "-o", obj]))

    def run(job):
        u, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (u, r.stdout, r.stderr))
        log = os.path.join(OBJ, u.replace(".cu", ".ptxas.log"))
        with open(log, "w") as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return u

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, u.replace(".cu", ".o")) for u in _sources()]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                       "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
