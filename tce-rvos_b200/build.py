"""In-tree build of the native code (no GPU needed; nvcc cross-compiles sm_100a).

Replaces the reference's models/ops/setup.py + make.sh (setup.py:36-47 refuses to build without a
visible GPU and passes no -gencode; this build needs neither a GPU nor setuptools).

Artefacts, all under ``tce-rvos_b200/_native/`` (git-ignored, shipped to the GPU box by gpurun):

  libmsda_b200.so                                  the C-ABI library (include/msda_b200.h)
  MultiScaleDeformableAttention_update.<abi>.so    torch binding with the reference's module name

Usage:  python tce-rvos_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
NATIVE = PKG_DIR / "_native"
INCLUDE = PKG_DIR.parent / "include"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo"]   # no --use_fast_math: IEEE division / expf
# The shared CUDA runtime: one cudart in the process next to torch's (which is loaded first and has the same
# soname), and no copy of the whole runtime's symbol table inside the shipped library.
LINK_FLAGS = ["--cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]

CAPI_SOURCES = ["msda_capi.cu"]
CAPI_DEPS = ["msda_kernels.cuh", "msda_prologue_host.inl", "msda_host_pipeline.inl", "msda_aux_kernels.cuh", "msda_aux_host.inl"]
TORCH_SOURCES = ["msda_torch.cpp"]
MODULE_NAME = "MultiScaleDeformableAttention_update"


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return cand


def _digest(paths, extra=()) -> str:
    h = hashlib.sha256()
    for p in paths:
        h.update(Path(p).read_bytes())
    for e in extra:
        h.update(str(e).encode())
    return h.hexdigest()


def _stale(target: Path, stamp_value: str) -> bool:
    stamp = target.with_suffix(target.suffix + ".stamp")
    return not (target.exists() and stamp.exists() and stamp.read_text() == stamp_value)


def _write_stamp(target: Path, stamp_value: str) -> None:
    target.with_suffix(target.suffix + ".stamp").write_text(stamp_value)


# what the last build_capi / build_torch_binding call did with each artefact: "compiled" or "reused" (the stamp next to
# the binary equals the digest of its sources, headers and flags, i.e. it was built from exactly these sources)
ACTIONS = {}


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(map(str, cmd)), flush=True)
    res = subprocess.run(list(map(str, cmd)), capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"build step failed: {' '.join(map(str, cmd[:3]))} ...")
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr)


def capi_path() -> Path:
    return NATIVE / "libmsda_b200.so"


def module_path() -> Path:
    return NATIVE / f"{MODULE_NAME}{sysconfig.get_config_var('EXT_SUFFIX')}"


def build_capi(force: bool = False, verbose: bool = False) -> Path:
    NATIVE.mkdir(exist_ok=True)
    out = capi_path()
    srcs = [CSRC / s for s in CAPI_SOURCES]
    deps = srcs + [CSRC / d for d in CAPI_DEPS] + [INCLUDE / "msda_b200.h"]
    stamp = _digest(deps, ARCH_FLAGS + NVCC_FLAGS + LINK_FLAGS)
    if force or _stale(out, stamp):
        cmd = [_nvcc(), *ARCH_FLAGS, *NVCC_FLAGS, *LINK_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", out, *srcs]
        _run(cmd, verbose)
        _write_stamp(out, stamp)
        ACTIONS[out.name] = "compiled" + (" (forced)" if force else "")
    else:
        ACTIONS[out.name] = f"reused: stamp {stamp[:12]} matches the digest of the sources, headers and flags"
    return out


def build_torch_binding(force: bool = False, verbose: bool = False) -> Path:
    import torch
    from torch.utils import cpp_extension as ce

    NATIVE.mkdir(exist_ok=True)
    out = module_path()
    srcs = [CSRC / s for s in TORCH_SOURCES]
    stamp = _digest(srcs + [INCLUDE / "msda_b200.h"], [torch.__version__])
    if force or _stale(out, stamp):
        torch_lib = Path(torch.__file__).parent / "lib"
        cuda_home = Path(ce.CUDA_HOME or "/usr/local/cuda")
        inc = []
        for p in ce.include_paths("cuda"):
            inc += ["-isystem", p]
        inc += ["-isystem", sysconfig.get_paths()["include"]]
        abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w",
               f"-D_GLIBCXX_USE_CXX11_ABI={abi}", f"-DTORCH_EXTENSION_NAME={MODULE_NAME}",
               "-DTORCH_API_INCLUDE_EXTENSION_H", *inc, *srcs, "-o", out,
               f"-L{torch_lib}", f"-L{cuda_home / 'lib64'}", f"-L{NATIVE}",
               "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
               "-lmsda_b200", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}", f"-Wl,-rpath,{cuda_home / 'lib64'}"]
        _run(cmd, verbose)
        _write_stamp(out, stamp)
        ACTIONS[out.name] = "compiled" + (" (forced)" if force else "")
    else:
        ACTIONS[out.name] = f"reused: stamp {stamp[:12]} matches the digest of the sources, headers and flags"
    return out


def build_variant(name: str, defines, verbose: bool = False) -> Path:
    """Tuning aid: the same sources compiled with extra -D flags into _native_variants/<name>/
    (the torch binding is copied next to it; its rpath is $ORIGIN)."""
    build_all(False, verbose)
    vdir = PKG_DIR / "_native_variants" / name
    vdir.mkdir(parents=True, exist_ok=True)
    out = vdir / "libmsda_b200.so"
    srcs = [CSRC / s for s in CAPI_SOURCES]
    cmd = [_nvcc(), *ARCH_FLAGS, *NVCC_FLAGS, *LINK_FLAGS, *[f"-D{d}" for d in defines], "-shared", "-Xcompiler", "-fPIC",
           "-o", out, *srcs]
    _run(cmd, verbose)
    shutil.copy2(module_path(), vdir / module_path().name)
    return vdir


def build_all(force: bool = False, verbose: bool = False):
    return build_capi(force, verbose), build_torch_binding(force, verbose)


if __name__ == "__main__":
    a, b = build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(a)
    print(b)
