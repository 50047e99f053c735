"""Generates tests/golden/*.npz by running the REFERENCE's own ms_deform_attn_core_pytorch
(models/ops/functions/ms_deform_attn_func.py:67-87), imported unmodified from /root/reference, in
fp64 (forward + torch autograd backward) on seeded inputs.  Run in the build container only
(`python oracle/gen_golden.py`); the vectors are committed because /root/reference does not exist
on the GPU box.

The reference module does `import MultiScaleDeformableAttention_update` at import time
(ms_deform_attn_func.py:18); a placeholder module of that name is registered first — only the
pure-PyTorch function is used here, no extension code runs.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import msda_oracle as O  # noqa: E402

REF_ROOT = Path(os.environ.get("TCE_RVOS_REF", "/root/reference"))
GOLDEN = ROOT / "tests" / "golden"


def load_reference_function():
    path = REF_ROOT / "models" / "ops" / "functions" / "ms_deform_attn_func.py"
    if not path.exists():
        raise SystemExit(f"reference not found at {path}")
    sys.modules.setdefault("MultiScaleDeformableAttention_update",
                           types.ModuleType("MultiScaleDeformableAttention_update"))
    spec = importlib.util.spec_from_file_location("_ref_ms_deform_attn_func", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


# name -> (kind, N, levels, Lq, M, D, P, seed)
CASES = {
    # the reference's own test geometry (models/ops/test.py:21-25), CPU-seeded
    "testpy_tiny":      ("testpy", 1, O.LEVELS_TEST_PY, 2, 2, 2, 2, 3),
    "testpy_d32":       ("testpy", 2, O.LEVELS_TEST_PY, 3, 2, 32, 2, 4),
    "testpy_d71":       ("testpy", 1, O.LEVELS_TEST_PY, 2, 2, 71, 2, 5),
    # TCE head geometry (8 heads x 32 ch, 4 levels x 4 points) on small maps
    "tce_small_model":  ("model", 1, [(6, 10), (3, 5), (2, 3), (1, 2)], None, 8, 32, 4, 0),
    "tce_small_stress": ("stress", 2, [(6, 10), (3, 5), (2, 3), (1, 2)], 10, 8, 32, 4, 1),
    "tce_decoder_lq5":  ("model", 2, [(6, 10), (3, 5), (2, 3), (1, 2)], 5, 8, 32, 4, 2),
    # other (L, P) splits of 16 and a non-fast-path head count
    "l2p8":             ("stress", 1, [(7, 9), (4, 5)], 6, 8, 32, 8, 6),
    "m4_l4p4":          ("stress", 1, [(6, 7), (3, 4), (2, 2), (1, 1)], 9, 4, 32, 4, 7),
    "m3_d16_l3p2":      ("stress", 2, [(5, 6), (3, 3), (2, 2)], 7, 3, 16, 2, 8),
}


def main():
    ref_fn = load_reference_function()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    for name, (kind, N, levels, Lq, M, D, P, seed) in CASES.items():
        inp = O.make_inputs(kind, N, levels, Lq=Lq, M=M, D=D, P=P, seed=seed, dtype=torch.float64)
        v = inp["value"].clone().requires_grad_(True)
        l = inp["loc"].clone().requires_grad_(True)
        a = inp["aw"].clone().requires_grad_(True)
        out = ref_fn(v, inp["shapes"], l, a)
        gv, gl, ga = torch.autograd.grad(out, (v, l, a), inp["grad_out"])
        np.savez_compressed(
            GOLDEN / f"{name}.npz",
            value=inp["value"].numpy().astype(np.float32 if D > 40 else np.float64),
            shapes=inp["shapes"].numpy(), lsi=inp["lsi"].numpy(),
            loc=inp["loc"].numpy(), aw=inp["aw"].numpy(), grad_out=inp["grad_out"].numpy(),
            out=out.detach().numpy(), grad_value=gv.numpy(), grad_loc=gl.numpy(), grad_aw=ga.numpy(),
            meta=np.array([N, Lq if Lq is not None else sum(h * w for h, w in levels), M, D, len(levels), P, seed]),
            kind=np.array(kind))
        print(f"{name}: out {tuple(out.shape)}  |out|max {out.abs().max():.4g}")
    sz = sum(p.stat().st_size for p in GOLDEN.glob("*.npz"))
    print(f"wrote {len(CASES)} cases, {sz / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
