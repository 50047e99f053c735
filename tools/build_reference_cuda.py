"""Builds the REFERENCE's own CUDA extension for sm_100 as a comparison point (never as product).

Reads the sources where they lie under /root/reference/models/ops/src (build container only),
copies them to a scratch dir under /tmp, applies the two-token fix torch >= 2.x needs
(`value.type()` -> `value.scalar_type()` at ms_deform_attn_cuda.cu:82,167), compiles
with TORCH_CUDA_ARCH_LIST=10.0 (no GPU needed) and writes ONLY the resulting shared object to
baseline/_ref/ (git-ignored; it travels to the GPU box with gpurun).  No reference source enters
the repository.  The module is named `MSDA_reference_cuda` so that it can be loaded next to the
B200 build of `MultiScaleDeformableAttention_update`.

Validity: D = 32, 2-D, int32-safe sizes only (its D >= 64 backward is defective and
its host code overflows int32 from N = 960 frames at 512x896).
"""
import glob
import os
import re
import shutil
import sys
import tempfile

REF = os.environ.get("TCE_RVOS_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "baseline", "_ref")
NAME = "MSDA_reference_cuda"


def main():
    src = os.path.join(REF, "models", "ops", "src")
    if not os.path.isdir(src):
        print(f"reference sources not found at {src}; nothing built")
        return 1
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils import cpp_extension as ce
    work = tempfile.mkdtemp(prefix="msda_ref_build_")
    dst = os.path.join(work, "src")
    shutil.copytree(src, dst)
    cu = os.path.join(dst, "cuda", "ms_deform_attn_cuda.cu")
    text = open(cu).read()
    text, n = re.subn(r"AT_DISPATCH_FLOATING_TYPES\(value\.type\(\)", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type()", text)
    assert n == 2, n
    open(cu, "w").write(text)
    sources = [os.path.join(dst, "vision.cpp")] + glob.glob(os.path.join(dst, "cpu", "*.cpp")) + glob.glob(os.path.join(dst, "cuda", "*.cu"))
    build_dir = os.path.join(work, "build")
    os.makedirs(build_dir)
    ce.load(name=NAME, sources=sources, extra_include_paths=[dst], extra_cflags=["-DWITH_CUDA", "-w"],
            extra_cuda_cflags=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                               "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__", "-w"],
            with_cuda=True, build_directory=build_dir, verbose=False, is_python_module=False)
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(build_dir, NAME + ".so")
    shutil.copy2(so, os.path.join(OUT, NAME + ".so"))
    shutil.rmtree(work, ignore_errors=True)
    print(os.path.join(OUT, NAME + ".so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
