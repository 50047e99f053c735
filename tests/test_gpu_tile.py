"""Parity of the persistent SM-resident kernels (csrc/msda_tile_kernels.cuh) — `-m gpu`, through the C ABI.

The persistent forward / backward replace the one-shot kernels for encoder-sized calls.  They must agree with
the fp64 oracle to the same tolerances and with the one-shot kernels to rounding, for every way the
schedule can unfold: all levels resident in shared memory, only the coarse ones, none; fewer units than
SMs (segments inside a unit), more units than SMs (claiming and helping); query counts that do not divide
the grab batch; gated-off and non-finite sampling points.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from test_gpu_parity import DEV, FWD_TOL, GRAD_TOL, BF16_TOL, check_grads, dev_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tile_forced(capi):
    lib = capi.load()
    prev = lib.msda_tile_mode(1)
    yield lib
    lib.msda_tile_mode(prev)


def _one_shot(capi, lib, t):
    prev = lib.msda_tile_mode(0)
    try:
        out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
        grads = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    finally:
        lib.msda_tile_mode(prev)
    return out, grads


def _kernel_names(lib, capi, t):
    sh = capi.shape_of(t["value"], t["loc"], t["shapes"].shape[0])
    return lib.msda_kernel_name(C.byref(sh), 0), lib.msda_kernel_name(C.byref(sh), 1)


CASES = [
    # (id, kind, frames, levels, Lq)
    ("all_levels_resident", "model", 3, [(12, 20), (6, 10), (3, 5), (2, 3)], None),      # S = 321 <= capacity
    ("coarse_resident_A", "model", 2, "A", None),                                        # levels 2-3 of 360x640
    ("coarse_resident_A_uniform", "testpy", 2, "A", 1000),
    ("only_last_level_C", "model", 1, "C", 3001),                                        # 16x28 + 8x14 does not fit
    ("more_units_than_sms", "model", 90, [(12, 20), (6, 10), (3, 5), (2, 3)], 37),       # 180 units: claim + help
    ("ragged_batches", "model", 5, [(10, 16), (5, 8), (3, 4), (2, 2)], 13),
    ("stress_locations", "stress", 4, "B", 777),
]


def _levels(oracle, spec):
    return {"A": oracle.LEVELS_A, "B": oracle.LEVELS_B, "C": oracle.LEVELS_C}[spec] if isinstance(spec, str) else spec


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tile_kernels_against_oracle_and_one_shot(capi, oracle, tile_forced, case):
    _, kind, N, lv, Lq = case
    levels = _levels(oracle, lv)
    inp = oracle.make_inputs(kind, N, levels, Lq=Lq, seed=11)
    t = dev_inputs(inp, torch.float32)
    assert _kernel_names(tile_forced, capi, t) == (b"fwd_tile_d32_f32", b"bwd_tile_d32_f32")
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    if kind == "stress":      # exact-integer coordinates: compare with the restatement of the kernels' own gating
        ref_out = oracle.c_forward(inp["value"], inp["shapes"], inp["lsi"], inp["loc"], inp["aw"], np.float64)
        rv, rl, ra = oracle.c_backward(inp["value"], inp["shapes"], inp["lsi"], inp["loc"], inp["aw"], inp["grad_out"], np.float64)
        ref_out = torch.as_tensor(ref_out).view(out.shape)
    else:
        ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(inp["value"].double(), inp["shapes"], inp["loc"].double(),
                                                          inp["aw"].double(), inp["grad_out"].double())
    assert oracle.max_rel(out.cpu(), ref_out) < FWD_TOL
    check_grads(oracle, got, (rv, rl, ra), inp["loc"], levels, GRAD_TOL, stress=kind == "stress")
    # the one-shot kernels do the same arithmetic per (query, head): forward and the two written-once
    # gradients agree to the last bits, grad_value up to summation order
    out1, (gv1, gl1, ga1) = _one_shot(capi, tile_forced, t)
    assert oracle.max_rel(out.cpu(), out1.cpu()) < 1e-6
    assert oracle.max_rel(got[1].cpu(), gl1.cpu()) < 1e-6 and oracle.max_rel(got[2].cpu(), ga1.cpu()) < 1e-6
    assert oracle.max_rel(got[0].cpu(), gv1.cpu()) < 1e-5


def test_tile_backward_is_repeatable_and_leaves_no_state(capi, oracle, tile_forced):
    """Two calls in a row (the workspace is scratch, re-zeroed by every call) give the same answer, and a
    frame's gradient does not depend on which other frames share the call (frames are independent)."""
    levels = oracle.LEVELS_A
    inp = oracle.make_inputs("model", 3, levels, seed=12)
    t = dev_inputs(inp, torch.float32)
    a = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    b = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert oracle.max_rel(a[0].cpu(), b[0].cpu()) < 1e-6
    one = {k: (v[1:2].contiguous() if k in ("value", "loc", "aw", "grad_out") else v) for k, v in t.items()}
    c = capi.backward(one["value"], one["shapes"], one["lsi"], one["loc"], one["aw"], one["grad_out"])
    assert torch.equal(c[1], a[1][1:2]) and torch.equal(c[2], a[2][1:2])
    assert oracle.max_rel(c[0].cpu(), a[0][1:2].cpu()) < 1e-6
    o3 = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    o1 = capi.forward(one["value"], one["shapes"], one["lsi"], one["loc"], one["aw"])
    assert torch.equal(o1, o3[1:2])


def test_tile_bf16_backward_with_fp32_scratch(capi, oracle, tile_forced):
    levels = oracle.LEVELS_A
    inp = oracle.make_inputs("model", 2, levels, seed=13)
    t = dev_inputs(inp, torch.bfloat16, torch.float32)
    assert _kernel_names(tile_forced, capi, t)[1] == b"bwd_tile_d32_bf16"
    got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    _, rv, rl, ra = oracle.core_pytorch_fwd_bwd(t["value"].cpu().double(), inp["shapes"], t["loc"].cpu().double(),
                                                t["aw"].cpu().double(), t["grad_out"].cpu().double())
    gv, gl, ga = (x.float() for x in got)
    check_grads(oracle, (gv, gl, ga), (rv, rl, ra), t["loc"].cpu(), levels, BF16_TOL)


def test_full_config2_runs_on_the_tile_kernels_by_default(capi, oracle):
    """BASELINE.json configs[1] (10 frames, 360x640): the dispatcher's own choice, no forcing."""
    lib = capi.load()
    assert lib.msda_tile_mode(-2) == -1            # query only
    levels = oracle.LEVELS_A
    inp = oracle.make_inputs("model", 10, levels, seed=0)
    t = dev_inputs(inp, torch.float32)
    assert _kernel_names(lib, capi, t) == (b"fwd_tile_d32_f32", b"bwd_tile_d32_f32")
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(inp["value"].double(), inp["shapes"], inp["loc"].double(),
                                                      inp["aw"].double(), inp["grad_out"].double())
    assert oracle.max_rel(out.cpu(), ref_out) < FWD_TOL
    check_grads(oracle, got, (rv, rl, ra), inp["loc"], levels, GRAD_TOL)


@pytest.mark.parametrize("forced", [0, 1])
def test_non_finite_and_overflowing_locations_contribute_nothing(capi, oracle, forced):
    """A sampling point whose location is inf / NaN / overflows when scaled is gated off by the reference
    (ms_deform_im2col_cuda.cuh:441: all four comparisons are false) — output and gradients of the other
    points must be unaffected, and the gated point's own gradients are zero (the fast kernels
    used to produce NaN through inf - inf)."""
    lib = capi.load()
    prev = lib.msda_tile_mode(forced)
    try:
        levels = [(12, 20), (6, 10), (3, 5), (2, 3)]
        inp = oracle.make_inputs("model", 2, levels, seed=14)
        loc = inp["loc"].clone()
        bad = torch.tensor([float("inf"), float("-inf"), float("nan"), 1e38, -1e38, 3.0e30])
        flat = loc.view(-1, 2)
        idx = torch.arange(0, flat.shape[0], 17)
        flat[idx, 0] = bad[torch.arange(idx.numel()) % bad.numel()]
        flat[idx[::2], 1] = bad[(torch.arange(idx[::2].numel()) + 2) % bad.numel()]
        t = dev_inputs(dict(inp, loc=loc), torch.float32)
        out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
        gv, gl, ga = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
        assert torch.isfinite(out).all() and torch.isfinite(gv).all() and torch.isfinite(ga).all()
        # truth: the same call with the bad points moved far outside (finite), where every implementation gates them
        good = loc.clone()
        nf = ~torch.isfinite(good) | (good.abs() > 1e6)
        good[nf] = 5.0
        ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(inp["value"].double(), inp["shapes"], good.double(),
                                                          inp["aw"].double(), inp["grad_out"].double())
        assert oracle.max_rel(out.cpu(), ref_out) < FWD_TOL
        assert oracle.max_rel(gv.cpu(), rv) < GRAD_TOL and oracle.max_rel(ga.cpu(), ra) < GRAD_TOL
        bad_pt = nf.any(-1)
        assert float(ga.cpu()[bad_pt].abs().max()) == 0.0
        gl_bad = gl.cpu()[bad_pt]
        assert float(torch.nan_to_num(gl_bad, nan=1.0).abs().max()) == 0.0
    finally:
        lib.msda_tile_mode(prev)
