"""C-ABI library and host-side logic, no GPU: the library loads, exports every symbol the header
declares, validates arguments before touching CUDA, and the Python mirror keeps the reference's
API."""
import ctypes as C
import re
import sys
import types
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__
    __graft_entry__.build()


def header_functions():
    text = (ROOT / "include" / "msda_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(capi):
    lib = C.CDLL(str(capi.library_path()))
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/msda_b200.h but not exported"
    assert set(names) == set(capi.PROTOTYPES), set(names) ^ set(capi.PROTOTYPES)


def test_introspection(capi):
    lib = capi.load()
    assert lib.msda_abi_version() == 1
    assert b"im2col_step" in lib.msda_error_string(capi.MSDA_E_IM2COL_STEP)
    sh = capi.make_shape(10, 4820, 4820, 8, 32, 4, 4, capi.MSDA_F32, capi.MSDA_F32)
    # SURVEY.md 8d worked example: fwd 172.7 MB, fwd+bwd 468.9 MB at config 2
    assert lib.msda_algorithmic_bytes(sh, 0) == 4 * 10 * (256 * 4820 + 640 * 4820)
    assert lib.msda_algorithmic_bytes(sh, 2) == 4 * 10 * (768 * 4820 + 1664 * 4820)
    assert lib.msda_kernel_name(sh, 0) == b"fwd_fast_d32_f32"
    assert lib.msda_kernel_name(sh, 1) == b"bwd_fast_d32_f32"
    sh16 = capi.make_shape(10, 4820, 4820, 8, 32, 4, 4, capi.MSDA_BF16, capi.MSDA_F32)
    assert lib.msda_kernel_name(sh16, 0) == b"fwd_fast_d32_bf16"
    assert lib.msda_backward_workspace_bytes(sh16) == 10 * 4820 * 256 * 4
    assert lib.msda_backward_workspace_bytes(sh) == 0
    odd = capi.make_shape(1, 30, 2, 2, 71, 2, 2, capi.MSDA_F64, capi.MSDA_F64)
    assert lib.msda_kernel_name(odd, 1) == b"bwd_generic"
    assert lib.msda_prologue_supported(sh, 2) == 1 and lib.msda_prologue_supported(sh, 3) == 0
    assert lib.msda_prologue_supported(odd, 2) == 0


def test_im2col_step_rule(capi):
    """ms_deform_attn_cuda.cu:54-56: step = min(N, im2col_step); N % step must be 0."""
    lib = capi.load()
    assert lib.msda_check_im2col_step(36, 64) == 0
    assert lib.msda_check_im2col_step(64, 64) == 0
    assert lib.msda_check_im2col_step(128, 64) == 0
    assert lib.msda_check_im2col_step(65, 64) == capi.MSDA_E_IM2COL_STEP
    assert lib.msda_check_im2col_step(100, 64) == capi.MSDA_E_IM2COL_STEP
    assert lib.msda_check_im2col_step(2, 2) == 0


def test_argument_validation_without_gpu(capi):
    lib = capi.load()
    sh = capi.make_shape(1, 12, 3, 8, 32, 1, 16, capi.MSDA_F32, capi.MSDA_F32)
    assert lib.msda_forward(C.byref(sh), None, None, None, None, None, None, None) == capi.MSDA_E_NULL
    assert lib.msda_forward(None, None, None, None, None, None, None, None) == capi.MSDA_E_NULL
    bad = capi.make_shape(1, 12, 3, 8, 0, 1, 16, capi.MSDA_F32, capi.MSDA_F32)
    assert lib.msda_forward(C.byref(bad), None, None, None, None, None, None, None) == capi.MSDA_E_SHAPE
    mixed = capi.make_shape(1, 12, 3, 8, 32, 1, 16, capi.MSDA_F32, capi.MSDA_BF16)
    assert lib.msda_forward(C.byref(mixed), None, None, None, None, None, None, None) == capi.MSDA_E_DTYPE
    empty = capi.make_shape(0, 12, 3, 8, 32, 1, 16, capi.MSDA_F32, capi.MSDA_F32)
    assert lib.msda_forward(C.byref(empty), None, None, None, None, None, None, None) == 0
    with pytest.raises(capi.MsdaError):
        capi.check(capi.MSDA_E_SHAPE, "x")


def test_cpu_tensors_are_rejected_like_the_reference(pkg):
    """models/ops/src/ms_deform_attn.h:39 / cpu/ms_deform_attn_cpu.cpp:26: no CPU implementation."""
    v = torch.zeros(1, 4, 8, 32)
    shapes = torch.tensor([[2, 2]])
    lsi = torch.tensor([0])
    loc = torch.zeros(1, 1, 8, 1, 16, 2)
    aw = torch.zeros(1, 1, 8, 1, 16)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        pkg.MSDeformAttnFunction.apply(v, shapes, lsi, loc, aw, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        pkg.load_extension().ms_deform_attn_backward(v, shapes, lsi, loc, aw, torch.zeros(1, 1, 256), 64, False)
    m = pkg.MSDeformAttn()
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        m(torch.zeros(1, 1, 256), torch.zeros(1, 1, 4, 2), torch.zeros(1, 4, 256),
          torch.tensor([[1, 1]] * 4), torch.tensor([0, 1, 2, 3]))


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: no file under tce-rvos_b200/ may reference it."""
    for p in (ROOT / "tce-rvos_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".inl", ".h") and "_native" not in p.parts:
            text = p.read_text()
            assert "oracle" not in text.replace("oracle/", "ORACLE_DOC_PATH"), f"{p} mentions the oracle"
            assert "grid_sample" not in text, f"{p} contains a PyTorch sampling fallback"


def test_module_mirrors_reference_api(pkg):
    m = pkg.MSDeformAttn(d_model=256, n_levels=4, n_heads=8, n_points=4)
    assert sorted(m.state_dict()) == sorted(
        f"{n}.{k}" for n in ("sampling_offsets", "attention_weights", "value_proj", "output_proj")
        for k in ("weight", "bias"))
    assert m.im2col_step == 64
    # offset bias = 8 compass directions x (p+1) px (ms_deform_attn.py:63-71)
    b = m.sampling_offsets.bias.view(8, 4, 4, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.tensor([1., 2., 3., 4.]))
    assert torch.allclose(b[2, 1, :, 1], torch.tensor([1., 2., 3., 4.]))
    assert float(m.attention_weights.weight.abs().max()) == 0.0
    with pytest.raises(ValueError):
        pkg.MSDeformAttn(d_model=250, n_heads=8)


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present on this machine")
def test_drop_in_under_the_reference_module_name(pkg):
    """The reference's own functions/modules import MultiScaleDeformableAttention_update
    (ms_deform_attn_func.py:18); with install_drop_in() that name resolves to the B200 build, the
    reference module constructs with identical parameters, and a reference state_dict loads."""
    import importlib.util
    pkg.install_drop_in()
    ext = pkg.load_extension()
    assert sys.modules["MultiScaleDeformableAttention_update"] is ext
    ops = REF / "models" / "ops"
    spec = importlib.util.spec_from_file_location("refops", ops / "__init__.py" if (ops / "__init__.py").exists() else ops / "functions" / "__init__.py",
                                                  submodule_search_locations=[str(ops)])
    refops = types.ModuleType("refops")
    refops.__path__ = [str(ops)]
    sys.modules["refops"] = refops
    import importlib
    ref_mod = importlib.import_module("refops.modules.ms_deform_attn")
    ref_fn = importlib.import_module("refops.functions.ms_deform_attn_func")
    assert ref_fn.MSDA is ext                                     # the reference now calls our kernels
    torch.manual_seed(0)
    theirs = ref_mod.MSDeformAttn(256, 4, 8, 4)
    ours = pkg.MSDeformAttn(256, 4, 8, 4)
    ours.load_state_dict(theirs.state_dict())                     # strict: same keys and shapes
    for k, v in theirs.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k])
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        theirs(torch.zeros(1, 1, 256), torch.zeros(1, 1, 4, 2), torch.zeros(1, 4, 256),
               torch.tensor([[1, 1]] * 4), torch.tensor([0, 1, 2, 3]))


def test_aux_entry_points_validate_before_touching_cuda(capi):
    """msda_value_epilogue / msda_topk_samples (rows f2, f3) reject bad arguments without a GPU."""
    lib = capi.load()
    F32, BF16, F64 = capi.MSDA_F32, capi.MSDA_BF16, capi.MSDA_F64
    assert lib.msda_value_epilogue(0, 256, None, F32, None, None, F32, None) == 0              # nothing to do
    assert lib.msda_value_epilogue(4, 256, None, F32, None, None, F32, None) == capi.MSDA_E_NULL
    assert lib.msda_value_epilogue(-1, 256, None, F32, None, None, F32, None) == capi.MSDA_E_SHAPE
    buf = (C.c_char * 4096)()
    addr = (C.addressof(buf) + 15) & ~15                                                        # 16-byte aligned host address
    assert lib.msda_value_epilogue(2, 12, addr, F32, None, addr, F32, None) == capi.MSDA_E_UNSUPPORTED   # width % 8
    assert lib.msda_value_epilogue(2, 16, addr + 4, F32, None, addr, F32, None) == capi.MSDA_E_ALIGN
    assert lib.msda_value_epilogue(2, 16, addr, F64, None, addr, F32, None) == capi.MSDA_E_DTYPE

    sh = capi.make_shape(2, 1, 5, 8, 1, 4, 4, F32, F32)
    assert lib.msda_topk_samples(C.byref(sh), 30, None, None, None, None, None, None, None) == capi.MSDA_E_NULL
    assert lib.msda_topk_samples(C.byref(sh), 0, addr, addr, None, addr, addr, None, None) == capi.MSDA_E_UNSUPPORTED
    assert lib.msda_topk_samples(C.byref(sh), 129, addr, addr, None, addr, addr, None, None) == capi.MSDA_E_UNSUPPORTED
    big = capi.make_shape(2, 1, 5, 8, 1, 4, 32, F32, F32)                                       # 1024 candidates
    assert lib.msda_topk_samples(C.byref(big), 30, addr, addr, None, addr, addr, None, None) == capi.MSDA_E_UNSUPPORTED
    empty = capi.make_shape(0, 1, 5, 8, 1, 4, 4, F32, F32)
    assert lib.msda_topk_samples(C.byref(empty), 30, None, None, None, None, None, None, None) == 0
    half = capi.make_shape(2, 1, 5, 8, 1, 4, 4, F32, capi.MSDA_F16)
    assert lib.msda_topk_samples(C.byref(half), 30, addr, addr, None, addr, addr, None, None) == capi.MSDA_E_DTYPE
    assert lib.msda_topk_samples(None, 30, addr, addr, None, addr, addr, None, None) == capi.MSDA_E_NULL


def test_header_is_plain_c_and_links(tmp_path, capi):
    """include/msda_b200.h is the drop-in boundary: it must compile as C99 (no C++ or torch types) and a
    plain C program must link against libmsda_b200.so and call it (introspection only: no GPU here)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use_abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "msda_b200.h"\n'
        "int main(void) {\n"
        "  msda_shape s = {10, 4820, 4820, 8, 32, 4, 4, MSDA_F32, MSDA_F32};\n"
        '  printf("%d %s %lld %d\\n", msda_abi_version(), msda_kernel_name(&s, 1),\n'
        "         (long long)msda_algorithmic_bytes(&s, 2), msda_forward(&s, 0, 0, 0, 0, 0, 0, 0));\n"
        "  return 0;\n}\n")
    exe = tmp_path / "use_abi"
    lib = capi.library_path()
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
                    f"-L{lib.parent}", "-lmsda_b200", f"-Wl,-rpath,{lib.parent}", "-Wl,-rpath,/usr/local/cuda/lib64"],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["1", "bwd_fast_d32_f32", str(4 * 10 * (768 * 4820 + 1664 * 4820)), str(capi.MSDA_E_NULL)]


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present on this machine")
def test_package_replaces_the_models_ops_directory(tmp_path):
    """The package mirrors the layout of the reference's models/ops (functions/, modules/), so the directory
    can be replaced wholesale: with models/ops pointing at tce-rvos_b200 the UNMODIFIED
    models/tce_deformable_transformer.py imports this repo's MSDeformAttn and builds its layers with it."""
    import os
    import subprocess
    (tmp_path / "models").mkdir()
    for p in (REF / "models").iterdir():
        if p.name != "ops":
            os.symlink(p, tmp_path / "models" / p.name)
    os.symlink(ROOT / "tce-rvos_b200", tmp_path / "models" / "ops")
    os.symlink(REF / "util", tmp_path / "util")
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, '.')\n"
        "import models.tce_deformable_transformer as T\n"
        "from models.ops.modules import MSDeformAttn\n"
        "from models.ops.functions import MSDeformAttnFunction\n"
        "layer = T.DeformableTransformerEncoderLayer(256, 1024, 0.1, 'relu', 4, 8, 4)\n"
        "assert type(layer.self_attn) is MSDeformAttn and hasattr(layer.self_attn, 'fused_prologue')\n"
        "print('ok', MSDeformAttn.__module__)\n" % str(ROOT / "tools" / "shims"))
    res = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    assert res.stdout.split()[-2:] == ["ok", "models.ops.modules.ms_deform_attn"]
