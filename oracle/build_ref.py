"""TEST INFRASTRUCTURE ONLY. Compile the reference's own native ops into oracle/_ref/.

Builds `vision3d/ops/csrc` of the read-only reference tree as torch extensions, from the
sources where they lie. The reference was written for torch 1.4; torch 2.11 needs ONE token
changed in two places (`dets.type()` -> `dets.scalar_type()` inside the AT_DISPATCH macros at
nms_rotated_cpu.cpp:67 and nms_rotated_cuda.cu:98). The patch is applied on the fly to a
scratch copy under a temp dir; no reference source enters the repo, only the built .so files
land in oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).

  oracle/_ref/ref_C_cpu.so    vision.cpp + *_cpu.cpp             -> `kind: "reference"` CPU baseline
  oracle/_ref/ref_C_cuda.so   + *_cuda.cu for sm_100a (WITH_CUDA) -> the recompiled reference
                              SIMT kernels, used as the on-GPU comparator in tests/bench.

Do not call get_compiler_version() from these modules (std::ostringstream segfaults with this
image's g++ wrapper).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("V3D_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def _scratch_sources():
    src = os.path.join(REF, "vision3d", "ops", "csrc")
    tmp = tempfile.mkdtemp(prefix="v3d_refsrc_")
    dst = os.path.join(tmp, "csrc")
    shutil.copytree(src, dst)
    subprocess.check_call(["chmod", "-R", "u+w", dst])
    for rel, old, new in [
        ("nms_rotated/nms_rotated_cpu.cpp", "AT_DISPATCH_FLOATING_TYPES(dets.type()",
         "AT_DISPATCH_FLOATING_TYPES(dets.scalar_type()"),
        ("nms_rotated/nms_rotated_cuda.cu", "dets_sorted.type(), \"nms_rotated_kernel_cuda\"",
         "dets_sorted.scalar_type(), \"nms_rotated_kernel_cuda\""),
    ]:
        p = os.path.join(dst, rel)
        s = open(p).read()
        assert old in s, (rel, "token to patch not found")
        open(p, "w").write(s.replace(old, new))
    return tmp, dst


def build(with_cuda=True, verbose=False):
    if not os.path.isdir(REF):
        print("reference tree not mounted; keeping prebuilt oracle/_ref", file=sys.stderr)
        return
    os.makedirs(OUT, exist_ok=True)
    from torch.utils.cpp_extension import load
    tmp, csrc = _scratch_sources()
    try:
        cpu_src = [os.path.join(csrc, "vision.cpp"),
                   os.path.join(csrc, "box_iou_rotated", "box_iou_rotated_cpu.cpp"),
                   os.path.join(csrc, "nms_rotated", "nms_rotated_cpu.cpp")]
        bdir = os.path.join(tmp, "b_cpu")
        os.makedirs(bdir)
        load(name="ref_C_cpu", sources=cpu_src, extra_include_paths=[csrc],
             extra_cflags=["-O3", "-w"], build_directory=bdir, is_python_module=False,
             verbose=verbose)
        shutil.copy(os.path.join(bdir, "ref_C_cpu.so"), os.path.join(OUT, "ref_C_cpu.so"))
        if with_cuda:
            cu_src = cpu_src + [os.path.join(csrc, "box_iou_rotated", "box_iou_rotated_cuda.cu"),
                                os.path.join(csrc, "nms_rotated", "nms_rotated_cuda.cu"),
                                os.path.join(csrc, "cuda_version.cu")]
            bdir = os.path.join(tmp, "b_cuda")
            os.makedirs(bdir)
            os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
            load(name="ref_C_cuda", sources=cu_src, extra_include_paths=[csrc],
                 extra_cflags=["-O3", "-w", "-DWITH_CUDA"],
                 extra_cuda_cflags=["-O3", "-w", "-DWITH_CUDA",
                                    "-gencode", "arch=compute_100a,code=sm_100a"],
                 build_directory=bdir, is_python_module=False, with_cuda=True, verbose=verbose)
            shutil.copy(os.path.join(bdir, "ref_C_cuda.so"), os.path.join(OUT, "ref_C_cuda.so"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    build(with_cuda="--no-cuda" not in sys.argv, verbose="-v" in sys.argv)
    print(sorted(os.listdir(OUT)))
