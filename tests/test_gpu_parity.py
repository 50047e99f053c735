"""Parity of the CUDA path with the oracle — the tests proper (run with `-m gpu` on a B200).

Everything here goes through the C ABI: either directly (capi.forward / capi.backward = ctypes on
libmsda_b200.so) or through the torch binding that forwards to the same entry points.

Tolerances (BASELINE.json north_star): fp32 forward within 1e-5 of an fp64 run of the reference
algorithm, gradients within 1e-4, bf16 within 1e-2; errors are measured relative to the tensor's
scale (max |a-b| / max |b|), and grad_sampling_loc excludes points within 1e-4 px (fp32; 2e-2 px
for bf16 locations) of a bilinear cell boundary, where the gradient is discontinuous.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
FWD_TOL, GRAD_TOL, BF16_TOL = 1e-5, 1e-4, 1e-2


def to_dev(g, dtype, keys=("value", "loc", "aw", "grad_out")):
    t = {k: torch.from_numpy(np.asarray(g[k])).to(DEV, dtype).contiguous() for k in keys}
    t["shapes"] = torch.from_numpy(np.asarray(g["shapes"])).to(DEV)
    t["lsi"] = torch.from_numpy(np.asarray(g["lsi"])).to(DEV)
    return t


def dev_inputs(inp, dtype, loc_dtype=None):
    loc_dtype = loc_dtype or dtype
    return dict(value=inp["value"].to(DEV, dtype), loc=inp["loc"].to(DEV, loc_dtype), aw=inp["aw"].to(DEV, loc_dtype),
                grad_out=inp["grad_out"].to(DEV, dtype), shapes=inp["shapes"].to(DEV), lsi=inp["lsi"].to(DEV))


def check_grads(O, got, ref, loc, levels, tol, eps_px=1e-4, stress=False):
    gv, gl, ga = got
    rv, rl, ra = ref
    O.assert_close(gv.cpu(), rv, tol, "grad_value")        # |a-b| <= tol*max|b| elementwise (scale-relative, see oracle)
    O.assert_close(ga.cpu(), ra, tol, "grad_attn")
    err, l2, frac = O.grad_loc_errors(gl.cpu(), rl, loc, levels, eps_px)
    assert err < tol, ("grad_loc", err, l2, frac)
    if not stress:
        assert l2 < 10 * tol, ("grad_loc rel-L2", l2, frac)


# ------------------------------------------------------------------------------------------------
# golden vectors of the reference (every case, fp64 generic kernels and fp32 fast/generic kernels)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_golden_fp64_through_c_abi(capi, oracle, name):
    g = load_golden(name)
    t = to_dev(g, torch.float64)
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    assert oracle.max_rel(out.cpu().view(g["out"].shape), g["out"]) < 1e-12
    got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    check_grads(oracle, got, (g["grad_value"], g["grad_loc"], g["grad_aw"]), g["loc"], g["shapes"].tolist(),
                1e-11, eps_px=1e-9, stress=str(g["kind"]) == "stress")


@pytest.mark.parametrize("name", golden_names())
def test_golden_fp32_through_c_abi(capi, oracle, name):
    g = load_golden(name)
    t = to_dev(g, torch.float32)
    # the fp64 truth for fp32-rounded inputs is the golden itself up to input rounding (1e-7)
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(t["value"].cpu().double(), g["shapes"], t["loc"].cpu().double(),
                                                      t["aw"].cpu().double(), t["grad_out"].cpu().double())
    assert oracle.max_rel(out.cpu(), ref_out) < FWD_TOL
    got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    check_grads(oracle, got, (rv, rl, ra), t["loc"].cpu(), g["shapes"].tolist(), GRAD_TOL,
                stress=str(g["kind"]) == "stress")


def test_fast_kernels_are_selected_for_tce_geometry(capi):
    lib = capi.load()
    sh = capi.make_shape(2, 83, 83, 8, 32, 4, 4, capi.MSDA_F32, capi.MSDA_F32)
    assert lib.msda_kernel_name(C.byref(sh), 0) == b"fwd_fast_d32_f32"


# ------------------------------------------------------------------------------------------------
# torch binding == C ABI, autograd plumbing, reference test.py checks
# ------------------------------------------------------------------------------------------------
def test_autograd_function_matches_c_abi(pkg, capi, oracle):
    inp = oracle.make_inputs("model", 2, [(12, 20), (6, 10), (3, 5), (2, 3)], seed=1)
    t = dev_inputs(inp, torch.float32)
    v, l, a = (t[k].clone().requires_grad_(True) for k in ("value", "loc", "aw"))
    out = pkg.MSDeformAttnFunction.apply(v, t["shapes"], t["lsi"], l, a, 64)
    out.backward(t["grad_out"])
    out2 = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    assert torch.equal(out, out2)                       # same kernel, deterministic forward
    gv, gl, ga = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    assert torch.equal(l.grad, gl) and torch.equal(a.grad, ga)      # written once, no atomics
    assert oracle.max_rel(v.grad.cpu(), gv.cpu()) < 1e-6            # atomics: order may differ


def test_reference_test_py_forward_and_gradcheck(pkg, oracle):
    """models/ops/test.py: forward double (:31-44), forward float (:47-60), gradcheck in fp64 for
    D in {30,32,64,71,1025,2048,3096} (:63-78,:85-86)."""
    levels = oracle.LEVELS_TEST_PY
    inp = oracle.make_inputs("testpy", 1, levels, Lq=2, M=2, D=2, P=2, seed=3, dtype=torch.float64)
    t = dev_inputs(inp, torch.float64)
    ref = oracle.core_pytorch(inp["value"], inp["shapes"], inp["loc"], inp["aw"])
    out = pkg.MSDeformAttnFunction.apply(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], 2)
    assert torch.allclose(out.cpu(), ref)
    out32 = pkg.MSDeformAttnFunction.apply(t["value"].float(), t["shapes"], t["lsi"], t["loc"].float(), t["aw"].float(), 2)
    assert torch.allclose(out32.cpu().double(), ref, rtol=1e-2, atol=1e-3)
    for D in (30, 32, 64, 71, 1025, 2048, 3096):
        inp = oracle.make_inputs("testpy", 1, levels, Lq=2, M=2, D=D, P=2, seed=D, dtype=torch.float64)
        t = dev_inputs(inp, torch.float64)
        v, l, a = (t[k].requires_grad_(True) for k in ("value", "loc", "aw"))
        assert torch.autograd.gradcheck(pkg.MSDeformAttnFunction.apply, (v, t["shapes"], t["lsi"], l, a, 2)), D


def test_preconditions_raise_like_the_reference(pkg, oracle):
    inp = oracle.make_inputs("model", 3, [(4, 4)], Lq=4, M=8, D=32, P=16, seed=0)
    t = dev_inputs(inp, torch.float32)
    f = pkg.MSDeformAttnFunction.apply
    with pytest.raises(RuntimeError, match="contiguous"):
        f(t["value"].transpose(0, 1).contiguous().transpose(0, 1), t["shapes"], t["lsi"], t["loc"], t["aw"], 64)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        f(t["value"], t["shapes"].cpu(), t["lsi"], t["loc"], t["aw"], 64)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        f(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], 2)       # 3 % 2 != 0 (.cu:54-56)
    assert f(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], 3).shape == (3, 4, 256)


def test_empty_inputs(pkg, capi, oracle):
    inp = oracle.make_inputs("model", 2, [(4, 4)], Lq=4, M=8, D=32, P=16, seed=0)
    t = dev_inputs(inp, torch.float32)
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"][:, :0].contiguous(), t["aw"][:, :0].contiguous())
    assert out.shape == (2, 0, 256)
    gv, gl, ga = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"][:, :0].contiguous(),
                               t["aw"][:, :0].contiguous(), t["grad_out"][:, :0].contiguous())
    assert float(gv.abs().max()) == 0.0 and gl.numel() == 0 and ga.numel() == 0
    out0 = capi.forward(t["value"][:0], t["shapes"], t["lsi"], t["loc"][:0], t["aw"][:0])
    assert out0.shape == (0, 4, 256)


# ------------------------------------------------------------------------------------------------
# BASELINE.json config 2 at full size: N = 2x5 frames, level-set A, Lq = S, fp32 and bf16
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["model", "testpy"])
def test_config2_fp32_against_fp64_oracle(capi, oracle, kind):
    levels = oracle.LEVELS_A
    inp = oracle.make_inputs(kind, 10, levels, seed=0)
    t = dev_inputs(inp, torch.float32)
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(inp["value"].double(), inp["shapes"], inp["loc"].double(),
                                                      inp["aw"].double(), inp["grad_out"].double())
    assert oracle.max_rel(out.cpu(), ref_out) < FWD_TOL
    check_grads(oracle, got, (rv, rl, ra), inp["loc"], levels, GRAD_TOL)


def test_baseline_config0_forward_on_gpu(pkg, oracle):
    """BASELINE.json configs[0] (5 frames, level set A, fp32 forward) through the reference-facing
    MSDeformAttnFunction: equal to the reference's CPU path within the fp32 forward tolerance."""
    inp = oracle.make_inputs("model", 5, oracle.LEVELS_A, seed=0)
    t = dev_inputs(inp, torch.float32)
    out = pkg.MSDeformAttnFunction.apply(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], 64)
    ref64 = oracle.core_pytorch(inp["value"].double(), inp["shapes"], inp["loc"].double(), inp["aw"].double())
    assert oracle.max_rel(out.cpu(), ref64) < FWD_TOL
    ref32 = oracle.core_pytorch(inp["value"], inp["shapes"], inp["loc"], inp["aw"])      # what the reference returns on CPU
    assert torch.allclose(out.cpu(), ref32, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("loc_dtype", [torch.float32, torch.bfloat16])
def test_config2_bf16_against_fp64_oracle(capi, oracle, loc_dtype):
    levels = oracle.LEVELS_A
    inp = oracle.make_inputs("model", 10, levels, seed=1)
    t = dev_inputs(inp, torch.bfloat16, loc_dtype)
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    # truth: fp64 run of the reference algorithm on the SAME (bf16-rounded) inputs
    ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(t["value"].cpu().double(), inp["shapes"], t["loc"].cpu().double(),
                                                      t["aw"].cpu().double(), t["grad_out"].cpu().double())
    assert oracle.max_rel(out.float().cpu(), ref_out) < BF16_TOL
    gv, gl, ga = (x.float() for x in got)
    eps = 1e-4 if loc_dtype == torch.float32 else 2e-2
    check_grads(oracle, (gv, gl, ga), (rv, rl, ra), t["loc"].float().cpu(), levels, BF16_TOL, eps_px=eps)
    # direct 16-bit vector atomics (no fp32 scratch): one rounding per contribution -> looser
    gv2, gl2, ga2 = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"], use_workspace=False)
    assert oracle.max_rel(gv2.float().cpu(), rv) < 1.5e-1
    # the two paths use different lane layouts (8 lanes x 4 channels with the fp32 scratch, 4 lanes x 8 channels
    # without), so the channel dot products are summed in a different order
    tol = 1e-5 if loc_dtype == torch.float32 else BF16_TOL
    assert oracle.max_rel(gl2.float().cpu(), got[1].float().cpu()) < tol
    assert oracle.max_rel(ga2.float().cpu(), got[2].float().cpu()) < tol


def test_decoder_and_ftf_shapes(capi, oracle):
    """Lq = 5 (decoder) and Lq = 8 (FTF) on a 36-frame clip (BASELINE.json config 3 geometry)."""
    levels = oracle.LEVELS_A
    for Lq, seed in ((5, 2), (8, 3)):
        inp = oracle.make_inputs("model", 36, levels, Lq=Lq, seed=seed)
        t = dev_inputs(inp, torch.float32)
        out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
        got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
        ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(inp["value"].double(), inp["shapes"], inp["loc"].double(),
                                                          inp["aw"].double(), inp["grad_out"].double())
        assert oracle.max_rel(out.cpu(), ref_out) < FWD_TOL
        check_grads(oracle, got, (rv, rl, ra), inp["loc"], levels, GRAD_TOL)


def test_stress_locations_fp32(capi, oracle):
    levels = oracle.LEVELS_B
    inp = oracle.make_inputs("stress", 5, levels, Lq=512, seed=4)
    t = dev_inputs(inp, torch.float32)
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    # the C oracle restates the kernels' own gating, so exact-integer coordinates are comparable
    ref_out = oracle.c_forward(inp["value"], inp["shapes"], inp["lsi"], inp["loc"], inp["aw"], np.float64)
    rv, rl, ra = oracle.c_backward(inp["value"], inp["shapes"], inp["lsi"], inp["loc"], inp["aw"], inp["grad_out"], np.float64)
    assert oracle.max_rel(out.cpu().view(ref_out.shape), ref_out) < FWD_TOL
    check_grads(oracle, got, (rv, rl, ra), inp["loc"], levels, GRAD_TOL, stress=True)


# ------------------------------------------------------------------------------------------------
# size-independent properties at the scaling-sweep size (BASELINE.json config 5: level-set C)
# ------------------------------------------------------------------------------------------------
def test_properties_at_sweep_size(capi, oracle):
    levels = oracle.LEVELS_C
    N = 64
    inp = oracle.make_inputs("model", 2, levels, seed=5)
    S = inp["value"].shape[1]
    g = torch.Generator(device=DEV).manual_seed(0)
    loc = inp["loc"].to(DEV).repeat(N // 2, 1, 1, 1, 1, 1).contiguous()
    aw = inp["aw"].to(DEV).repeat(N // 2, 1, 1, 1, 1).contiguous()
    shapes, lsi = inp["shapes"].to(DEV), inp["lsi"].to(DEV)
    v1 = torch.randn(N, S, 8, 32, device=DEV, generator=g)
    v2 = torch.randn(N, S, 8, 32, device=DEV, generator=g)
    o1 = capi.forward(v1, shapes, lsi, loc, aw)
    o2 = capi.forward(v2, shapes, lsi, loc, aw)
    o12 = capi.forward(v1 + v2, shapes, lsi, loc, aw)
    # linearity in value
    assert oracle.max_rel((o1 + o2).cpu(), o12.cpu()) < 1e-5
    # frame-permutation equivariance, bitwise: frames are independent (the sharding property)
    perm = torch.randperm(N, device=DEV, generator=g)
    op = capi.forward(v1[perm].contiguous(), shapes, lsi, loc[perm].contiguous(), aw[perm].contiguous())
    assert torch.equal(op, o1[perm])
    # partition of unity: constant value, all points strictly inside -> out == sum(attn) == 1
    loc_in = loc.clamp(0.07, 0.93)              # >= half a pixel inside even on the 8x14 level
    ones = torch.ones_like(v1)
    o = capi.forward(ones, shapes, lsi, loc_in, aw)
    assert float((o - 1).abs().max()) < 1e-5
    # checksum of the scatter: sum_S grad_value[n,:,m,c] == sum_q grad_out[n,q,m,c] (weights sum to 1)
    go = torch.randn(N, S, 256, device=DEV, generator=g)
    gv, gl, ga = capi.backward(ones, shapes, lsi, loc_in, aw, go)
    lhs = gv.double().sum(1).view(N, 256)
    rhs = go.double().sum(1)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-5
    # constant value field -> no location gradient, attention gradient = sum_c grad_out
    assert float(gl.abs().max()) < 1e-3 * float(go.abs().max())
    ga_ref = go.view(N, S, 8, 32).sum(-1)[..., None, None].expand_as(ga)
    assert oracle.max_rel(ga.cpu(), ga_ref.cpu()) < 1e-5


# ------------------------------------------------------------------------------------------------
# fused module prologue (row f1) and the MSDeformAttn module
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_prologue(pkg, oracle, ref_dim, dtype):
    levels = [(12, 20), (6, 10), (3, 5), (2, 3)]
    N, Lq, M, L, P = 3, 50, 8, 4, 4
    g = torch.Generator().manual_seed(7)
    S = sum(h * w for h, w in levels)
    value = torch.randn(N, S, M, 32, generator=g)
    offsets = torch.randn(N, Lq, M, L, P, 2, generator=g) * 2.0
    logits = torch.randn(N, Lq, M, L * P, generator=g)
    ref = torch.rand(N, Lq, L, ref_dim, generator=g)
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:] * 0.3 + 0.05
    go = torch.randn(N, Lq, M * 32, generator=g)
    shapes, lsi = oracle.level_tensors(levels)
    ldt = torch.float32
    dv, doff, dlog, dref, dgo = (x.to(DEV, d).contiguous().requires_grad_(r) for x, d, r in
                                 ((value, dtype, True), (offsets, ldt, True), (logits, ldt, True), (ref, ldt, True), (go, dtype, False)))
    out, loc, aw = pkg.MSDeformAttnPrologueFunction.apply(dv, shapes.to(DEV), lsi.to(DEV), doff, dlog, dref, True)
    out.backward(dgo)
    # fp64 truth on the same (rounded) inputs: torch prologue + reference algorithm + autograd
    tv, toff, tlog, tref = (x.detach().cpu().double().requires_grad_(True) for x in (dv, doff, dlog, dref))
    tloc, taw = oracle.prologue(toff, tlog, tref, levels)
    tout = oracle.core_pytorch(tv, shapes, tloc, taw)
    tout.backward(dgo.cpu().double())
    tol_f, tol_g = (FWD_TOL, GRAD_TOL) if dtype == torch.float32 else (BF16_TOL, BF16_TOL)
    assert oracle.max_rel(loc.cpu(), tloc.detach()) < 1e-6
    assert oracle.max_rel(aw.cpu(), taw.detach()) < 1e-6
    assert oracle.max_rel(out.float().cpu(), tout.detach()) < tol_f
    assert oracle.max_rel(dv.grad.float().cpu(), tv.grad) < tol_g
    assert oracle.max_rel(dlog.grad.cpu(), tlog.grad) < tol_g
    assert oracle.max_rel(dref.grad.cpu(), tref.grad) < 10 * tol_g      # sums ~128 cell-sensitive terms
    err, l2, frac = oracle.grad_loc_errors(doff.grad.cpu(), toff.grad, tloc.detach(), levels)
    assert err < tol_g, (err, l2, frac)
    # without materialisation the output is identical
    out2 = pkg.MSDeformAttnPrologueFunction.apply(dv.detach(), shapes.to(DEV), lsi.to(DEV), doff.detach(), dlog.detach(), dref.detach(), False)
    assert torch.equal(out2, out.detach())


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("ref_dim", [2, 4])
def test_module_forward_backward_against_oracle_composition(pkg, oracle, fused, ref_dim):
    levels = [(12, 20), (6, 10), (3, 5), (2, 3)]
    S = sum(h * w for h, w in levels)
    N, Lq = 2, 40
    torch.manual_seed(0)
    m = pkg.MSDeformAttn(256, 4, 8, 4)
    with torch.no_grad():                                    # random (not init) weights so every path matters
        m.sampling_offsets.weight.normal_(0, 0.02)
        m.attention_weights.weight.normal_(0, 0.1)
    m = m.to(DEV)
    m.fused_prologue = fused
    query = torch.randn(N, Lq, 256)
    src = torch.randn(N, S, 256)
    ref = torch.rand(N, Lq, 4, ref_dim) * (0.4 if ref_dim == 4 else 1.0) + 0.05
    mask = torch.zeros(N, S, dtype=torch.bool)
    mask[:, -7:] = True
    shapes, lsi = oracle.level_tensors(levels)
    q, s_ = query.to(DEV).requires_grad_(True), src.to(DEV).requires_grad_(True)
    out, loc, aw = m(q, ref.to(DEV), s_, shapes.to(DEV), lsi.to(DEV), mask.to(DEV))
    assert out.shape == (N, Lq, 256) and loc.shape == (N, Lq, 8, 4, 4, 2) and aw.shape == (N, Lq, 8, 4, 4)
    out.sum().backward()
    # oracle composition in fp64 with the same weights (ms_deform_attn.py:95-115)
    md = pkg.MSDeformAttn(256, 4, 8, 4).double()
    md.load_state_dict({k: v.double().cpu() for k, v in m.state_dict().items()})
    qd, sd = query.double().requires_grad_(True), src.double().requires_grad_(True)
    val = md.value_proj(sd).masked_fill(mask[..., None], 0.0).view(N, S, 8, 32)
    off = md.sampling_offsets(qd).view(N, Lq, 8, 4, 4, 2)
    log = md.attention_weights(qd).view(N, Lq, 8, 16)
    tloc, taw = oracle.prologue(off, log, ref.double(), levels)
    tout = md.output_proj(oracle.core_pytorch(val, shapes, tloc, taw))
    tout.sum().backward()
    assert oracle.max_rel(out.detach().cpu(), tout.detach()) < 5e-5     # includes fp32 cuBLAS Linear layers
    assert oracle.max_rel(loc.detach().cpu(), tloc.detach()) < 1e-5
    assert oracle.max_rel(aw.detach().cpu(), taw.detach()) < 1e-5
    assert oracle.max_rel(s_.grad.cpu(), sd.grad) < 1e-3
    assert oracle.rel_l2(q.grad.cpu(), qd.grad) < 1e-3
    if fused:
        m.return_sampling = False
        o2, l2_, a2 = m(q.detach(), ref.to(DEV), s_.detach(), shapes.to(DEV), lsi.to(DEV), mask.to(DEV))
        assert l2_ is None and a2 is None and torch.equal(o2, out.detach())


# ------------------------------------------------------------------------------------------------
# the fork's 3-D variant (row f4)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol_f,tol_g", [(torch.float64, 1e-12, 1e-11), (torch.float32, FWD_TOL, GRAD_TOL)])
def test_3d_variant_against_c_oracle(pkg, oracle, dtype, tol_f, tol_g):
    levels = [(6, 7), (3, 4)]
    N, step = 4, 2                                            # two groups of two frames (im2col_step chunks)
    inp = oracle.make_inputs("stress", N, levels, Lq=9, M=4, D=32, P=2, seed=9, dtype=torch.float64, three_d=True)
    t = dev_inputs(inp, dtype)
    v, l, a = (t[k].clone().requires_grad_(True) for k in ("value", "loc", "aw"))
    out = pkg.MSDeformAttnFunction3D.apply(v, t["shapes"], t["lsi"], l, a, step)
    out.backward(t["grad_out"])
    src = {k: t[k].cpu().double() for k in ("value", "loc", "aw", "grad_out")}
    ref_out = oracle.c_forward(src["value"], inp["shapes"], inp["lsi"], src["loc"], src["aw"], np.float64, group=step)
    rv, rl, ra = oracle.c_backward(src["value"], inp["shapes"], inp["lsi"], src["loc"], src["aw"], src["grad_out"],
                                   np.float64, group=step)
    assert oracle.max_rel(out.detach().cpu().view(ref_out.shape), ref_out) < tol_f
    assert oracle.max_rel(v.grad.cpu(), rv) < tol_g
    assert oracle.max_rel(a.grad.cpu(), ra) < tol_g
    eps = 1e-9 if dtype == torch.float64 else 1e-4
    mask = oracle.boundary_mask(src["loc"], levels, eps)
    fi = src["loc"][..., 2] * step - 0.5
    mask |= ((fi - fi.floor()) < eps) | ((fi - fi.floor()) > 1 - eps)
    keep = (~mask)[..., None].expand_as(l.grad).cpu()
    scale = float(np.abs(rl).max())
    assert float(((l.grad.cpu().double() - torch.from_numpy(rl)).abs() * keep).max()) / scale < tol_g


# ------------------------------------------------------------------------------------------------
# host-buffer API, streams, CUDA graphs, multi-GPU frame sharding
# ------------------------------------------------------------------------------------------------
def test_host_buffer_api_matches_device_api(capi, oracle):
    levels = oracle.LEVELS_A
    inp = oracle.make_inputs("model", 5, levels, Lq=600, seed=6)
    t = dev_inputs(inp, torch.float32)
    out_d = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    gv_d, gl_d, ga_d = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    sh = capi.shape_of(inp["value"], inp["loc"], len(levels))
    ctx = capi.HostContext(0)
    pin = {k: inp[k].pin_memory() for k in ("value", "loc", "aw", "grad_out")}
    out = torch.empty(5, 600, 256).pin_memory()
    gv, gl, ga = (torch.empty_like(inp[k]).pin_memory() for k in ("value", "loc", "aw"))
    for _ in range(2):                                          # second call reuses the staging
        ctx.forward_backward(sh, pin["value"], inp["shapes"], inp["lsi"], pin["loc"], pin["aw"], pin["grad_out"],
                             out, gv, gl, ga)
    assert torch.equal(out, out_d.cpu()) and torch.equal(gl, gl_d.cpu()) and torch.equal(ga, ga_d.cpu())
    assert oracle.max_rel(gv, gv_d.cpu()) < 1e-6
    out2 = np.empty((5, 600, 256), dtype=np.float32)            # pageable numpy buffers work too
    ctx.forward(sh, inp["value"].numpy(), inp["shapes"].numpy(), inp["lsi"].numpy(), inp["loc"].numpy(),
                inp["aw"].numpy(), out2)
    assert np.array_equal(out2, out_d.cpu().numpy())
    ctx.close()


def test_runs_on_current_stream_and_in_cuda_graph(pkg, capi, oracle):
    inp = oracle.make_inputs("model", 4, [(12, 20), (6, 10), (3, 5), (2, 3)], seed=8)
    t = dev_inputs(inp, torch.float32)
    expect = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    egv, egl, ega = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        # warm-up on the side stream, then capture: needs a launch path with no sync and no allocation
        ext = pkg.load_extension()
        for _ in range(2):
            o = ext.ms_deform_attn_forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], 64, False)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            o = ext.ms_deform_attn_forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], 64, False)
            gv, gl, ga = ext.ms_deform_attn_backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"], 64, False)
        o.zero_(); gl.zero_(); ga.zero_()
        graph.replay()
    side.synchronize()
    assert torch.equal(o, expect) and torch.equal(gl, egl) and torch.equal(ga, ega)
    assert oracle.max_rel(gv.cpu(), egv.cpu()) < 1e-6


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_frame_sharding_across_two_gpus(pkg, oracle):
    import importlib
    sh = importlib.import_module("tce-rvos_b200.sharding")
    inp = oracle.make_inputs("model", 5, [(12, 20), (6, 10), (3, 5), (2, 3)], seed=10)
    t = dev_inputs(inp, torch.float32)
    whole = pkg.MSDeformAttnFunction.apply(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], 64)
    parts = []
    for r in range(2):
        d = f"cuda:{r}"
        v, l, a = (x.to(d).contiguous() for x in sh.shard_frames(r, 2, inp["value"], inp["loc"], inp["aw"]))
        parts.append(pkg.MSDeformAttnFunction.apply(v, inp["shapes"].to(d), inp["lsi"].to(d), l, a, 64).cpu())
    assert torch.equal(torch.cat(parts, 0), whole.cpu())


# ------------------------------------------------------------------------------------------------
# the reference's own CUDA extension, when its sm_100 build is present (tools/build_reference_cuda.py)
# ------------------------------------------------------------------------------------------------
def _reference_cuda():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "MSDA_reference_cuda.so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("MSDA_reference_cuda", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("kind", ["model", "stress"])
def test_against_reference_cuda_extension(pkg, oracle, kind):
    """Same inputs through the reference's CUDA kernels (fp32, D = 32) and through ours."""
    ref = _reference_cuda()
    if ref is None:
        pytest.skip("baseline/_ref/MSDA_reference_cuda.so not built")
    levels = oracle.LEVELS_A
    inp = oracle.make_inputs(kind, 5, levels, Lq=(None if kind == "model" else 700), seed=12)
    t = dev_inputs(inp, torch.float32)
    ext = pkg.load_extension()
    args = (t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    out_r = ref.ms_deform_attn_forward(*args, 64, False)
    out_o = ext.ms_deform_attn_forward(*args, 64, False)
    gr = ref.ms_deform_attn_backward(*args, t["grad_out"], 64, False)
    go = ext.ms_deform_attn_backward(*args, t["grad_out"], 64, False)
    torch.cuda.synchronize()
    assert oracle.max_rel(out_o.cpu(), out_r.cpu()) < 2e-6
    assert oracle.max_rel(go[0].cpu(), gr[0].cpu()) < 2e-6       # grad_value (atomic order differs)
    assert oracle.max_rel(go[2].cpu(), gr[2].cpu()) < 5e-6       # grad_attn_weight
    # both are fp32: cell selection is identical (same h_im/w_im arithmetic), so no boundary masking
    assert oracle.max_rel(go[1].cpu(), gr[1].cpu()) < 2e-5       # grad_sampling_loc


@pytest.mark.parametrize("kind,frames,step", [("model", 4, 2), ("testpy", 6, 3), ("testpy", 5, 64)])
def test_3d_variant_against_reference_cuda_extension(pkg, oracle, kind, frames, step):
    """Pins the fork's cross-frame variant (MSDeformAttnFunction3D, is_3d=True) to the reference itself: the
    reference has this path only as CUDA (ms_deform_im2col_cuda.cuh:366-420 forward, :565-712 backward), so its
    own kernels, compiled for sm_100 under baseline/_ref, are the ground truth here — forward and all three
    gradients, including the chunk-relative frame blend of the im2col_step loop (ms_deform_attn_cuda.cu:54,77)
    and a chunk of ONE group when step >= N."""
    ref = _reference_cuda()
    if ref is None:
        pytest.skip("baseline/_ref/MSDA_reference_cuda.so not built")
    levels = [(12, 20), (6, 10), (3, 5), (2, 3)]
    inp = oracle.make_inputs(kind, frames, levels, Lq=301, M=8, D=32, P=4, seed=21, three_d=True)
    t = dev_inputs(inp, torch.float32)
    ext = pkg.load_extension()
    args = (t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    out_r = ref.ms_deform_attn_forward(*args, step, True)
    out_o = ext.ms_deform_attn_forward(*args, step, True)
    gr = ref.ms_deform_attn_backward(*args, t["grad_out"], step, True)
    go = ext.ms_deform_attn_backward(*args, t["grad_out"], step, True)
    torch.cuda.synchronize()
    assert float(out_r.abs().max()) > 0
    assert oracle.max_rel(out_o.cpu(), out_r.cpu()) < 2e-6
    assert oracle.max_rel(go[0].cpu(), gr[0].cpu()) < 5e-6       # grad_value (atomic order differs)
    assert oracle.max_rel(go[2].cpu(), gr[2].cpu()) < 5e-6       # grad_attn_weight
    assert go[1].shape == gr[1].shape and go[1].shape[-1] == 3
    assert oracle.max_rel(go[1].cpu(), gr[1].cpu()) < 2e-5       # grad_sampling_loc incl. the frame coordinate


# ------------------------------------------------------------------------------------------------
# 64-bit addressing at the top of the scaling sweep (BASELINE.json config 5: N = 1024 at 512x896)
# ------------------------------------------------------------------------------------------------
def test_no_int32_overflow_at_1024_frames(capi, oracle):
    """The reference's host code computes chunk offsets in int32 and overflows from frame 896 on at
    this level set (ms_deform_attn_cuda.cu:62,84).  Here the element offset of the
    last frame is 1023*9520*256 = 2.49e9 > 2^31: its result must equal the same frame computed alone."""
    free, _ = torch.cuda.mem_get_info()
    if free < 90e9:
        pytest.skip("needs ~75 GB of free device memory")
    levels = oracle.LEVELS_C
    N, S, Lq = 1024, 9520, 9520
    one = oracle.make_inputs("model", 1, levels, seed=13)
    shapes, lsi = one["shapes"].to(DEV), one["lsi"].to(DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    value = torch.randn(N, S, 8, 32, device=DEV, generator=g)
    # every frame uses the same locations / weights except the probed ones, which get their own
    loc = one["loc"].to(DEV).expand(N, -1, -1, -1, -1, -1).contiguous()
    aw = one["aw"].to(DEV).expand(N, -1, -1, -1, -1).contiguous()
    go = torch.randn(N, Lq, 256, device=DEV, generator=g)
    probes = [0, 895, 896, 959, 1023]
    for i, n in enumerate(probes):
        loc[n] = (loc[n] + 0.003 * (i + 1)).clamp(-0.2, 1.2)
    out = capi.forward(value, shapes, lsi, loc, aw)
    gv, gl, ga = capi.backward(value, shapes, lsi, loc, aw, go)
    for n in probes:
        sl = slice(n, n + 1)
        o1 = capi.forward(value[sl].contiguous(), shapes, lsi, loc[sl].contiguous(), aw[sl].contiguous())
        assert torch.equal(o1, out[sl]), n
        gv1, gl1, ga1 = capi.backward(value[sl].contiguous(), shapes, lsi, loc[sl].contiguous(), aw[sl].contiguous(),
                                      go[sl].contiguous())
        assert torch.equal(gl1, gl[sl]) and torch.equal(ga1, ga[sl]), n
        assert oracle.max_rel(gv1.cpu(), gv[sl].cpu()) < 5e-6, n      # fp32 atomics: summation order differs
    # and one probed frame against the fp64 oracle
    n = 1023
    ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(value[n:n + 1].cpu().double(), one["shapes"], loc[n:n + 1].cpu().double(),
                                                      aw[n:n + 1].cpu().double(), go[n:n + 1].cpu().double())
    assert oracle.max_rel(out[n:n + 1].cpu(), ref_out) < FWD_TOL
    assert oracle.max_rel(gv[n:n + 1].cpu(), rv) < GRAD_TOL
    del value, loc, aw, go, out, gv, gl, ga
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------
# row f2: value epilogue (masked_fill + optional conversion), ms_deform_attn.py:95-98
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("src_dt,dst_dt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                           (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("width", [256, 24])
def test_value_epilogue_is_bit_exact(capi, oracle, src_dt, dst_dt, width):
    g = torch.Generator().manual_seed(3)
    N, S = 3, 4820 if width == 256 else 37
    v = torch.randn(N, S, width, generator=g).to(src_dt)
    v[0, 5, 3] = float("inf")                       # a masked row is never read: inf must not leak as NaN
    m = torch.rand(N, S, generator=g) < 0.25
    m[0, 5] = True
    ref = oracle.value_epilogue(v, m, dst_dt)
    got = capi.value_epilogue(v.to(DEV), m.to(DEV), dst_dt)
    assert got.dtype == dst_dt and torch.equal(got.cpu(), ref)
    # no mask: pure conversion / copy
    v[0, 5, 3] = 1.0
    assert torch.equal(capi.value_epilogue(v.to(DEV), None, dst_dt).cpu(), oracle.value_epilogue(v, None, dst_dt))
    if src_dt == dst_dt:                             # in place
        d = v.to(DEV)
        assert capi.value_epilogue(d, m.to(DEV), out=d).data_ptr() == d.data_ptr()
        assert torch.equal(d.cpu(), oracle.value_epilogue(v, m))
    assert capi.value_epilogue(v[:0].to(DEV), None, dst_dt).numel() == 0


def test_value_epilogue_autograd_and_module(pkg, oracle):
    g = torch.Generator().manual_seed(4)
    N, S = 2, 391
    x = torch.randn(N, S, 256, generator=g)
    m = torch.rand(N, S, generator=g) < 0.2
    go = torch.randn(N, S, 256, generator=g)
    for dt in (torch.float32, torch.bfloat16):
        a = x.clone().requires_grad_(True)
        oracle.value_epilogue(a, m, dt).backward(go.to(dt))
        b = x.to(DEV).requires_grad_(True)
        pkg.MSDeformAttnValueEpilogue.apply(b, m.to(DEV), dt).backward(go.to(DEV, dt))
        assert torch.equal(b.grad.cpu(), a.grad), dt
    # module: the fused pass gives the reference composition's result exactly in fp32 ...
    levels = [(12, 20), (6, 10), (3, 5), (2, 3)]
    S = sum(h * w for h, w in levels)
    shapes, lsi = oracle.level_tensors(levels, DEV)
    torch.manual_seed(0)
    mod = pkg.MSDeformAttn().to(DEV)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.01); mod.attention_weights.weight.normal_(0, 0.1)
    q = torch.randn(2, S, 256, generator=g).to(DEV)
    src = torch.randn(2, S, 256, generator=g).to(DEV)
    ref_pts = oracle.reference_points_grid(levels)[None, :, None, :].expand(2, S, 4, 2).contiguous().to(DEV)
    pm = (torch.rand(2, S, generator=g) < 0.15).to(DEV)
    base = mod(q, ref_pts, src, shapes, lsi, pm)[0]
    mod.fused_value_epilogue = True
    fused = mod(q, ref_pts, src, shapes, lsi, pm)[0]
    assert torch.equal(base, fused)
    # ... and within the bf16 tolerance with a 16-bit value
    mod.value_dtype = torch.bfloat16
    low = mod(q, ref_pts, src, shapes, lsi, pm)[0]
    assert low.dtype == torch.float32 and oracle.max_rel(low.cpu(), base.cpu()) < BF16_TOL
    low.sum().backward()
    assert mod.value_proj.weight.grad is not None and torch.isfinite(mod.value_proj.weight.grad).all()


# ------------------------------------------------------------------------------------------------
# row f3: decoder sample selection, tce_deformable_transformer.py:752-758
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,Lq,M,L,P,k", [(36, 5, 8, 4, 4, 30),      # TCE decoder at inference (36 frames, 5 queries)
                                          (10, 5, 8, 4, 4, 30),      # training clip
                                          (2, 301, 8, 4, 4, 30),
                                          (3, 9, 2, 4, 4, 30),       # 32 candidates
                                          (3, 9, 3, 4, 4, 7),        # 48 candidates: padded sort
                                          (2, 5, 8, 4, 16, 100),     # 512 candidates
                                          (2, 5, 8, 4, 4, 1), (2, 5, 8, 4, 4, 128)])
def test_decoder_sample_topk_is_bit_exact(pkg, capi, oracle, N, Lq, M, L, P, k):
    g = torch.Generator().manual_seed(N * 1000 + Lq)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.4 - 0.2
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    vr = 0.5 + 0.5 * torch.rand(N, L, 2, generator=g)
    rs, rw, ri = oracle.decoder_sample_topk(loc, aw, vr, k)
    s, w, i = capi.topk_samples(loc.to(DEV), aw.to(DEV), vr.to(DEV), k)
    assert torch.equal(i.cpu(), ri) and torch.equal(w.cpu(), rw) and torch.equal(s.cpu(), rs)
    s2, w2, i2 = pkg.decoder_sample_topk(loc.to(DEV), aw.to(DEV), vr.to(DEV), k)     # torch binding, same entry point
    assert torch.equal(i2, i) and torch.equal(w2, w) and torch.equal(s2, s)
    # no valid_ratios = no division
    s3, _, _ = capi.topk_samples(loc.to(DEV), aw.to(DEV), None, k)
    assert torch.equal(s3.cpu(), oracle.decoder_sample_topk(loc, aw, None, k)[0])


def test_decoder_sample_topk_ties_and_bf16(capi, oracle):
    g = torch.Generator().manual_seed(8)
    N, Lq = 4, 5
    loc = torch.rand(N, Lq, 8, 4, 4, 2, generator=g)
    vr = 0.5 + 0.5 * torch.rand(N, 4, 2, generator=g)
    # the model at initialisation: every weight is 1/16 (ms_deform_attn.py:72-73) -> the first 30 candidates
    flat = torch.full((N, Lq, 8, 4, 4), 1.0 / 16)
    s, w, i = capi.topk_samples(loc.to(DEV), flat.to(DEV), vr.to(DEV), 30)
    rs, rw, ri = oracle.decoder_sample_topk(loc, flat, vr, 30)
    assert torch.equal(i.cpu(), ri) and torch.equal(s.cpu(), rs) and torch.equal(w.cpu(), rw)
    # groups of equal weights, negative weights, and a bf16 run (many exact ties by construction)
    aw = (torch.randint(-3, 4, (N, Lq, 8, 4, 4), generator=g).float() / 8)
    s, w, i = capi.topk_samples(loc.to(DEV), aw.to(DEV), vr.to(DEV), 30)
    rs, rw, ri = oracle.decoder_sample_topk(loc, aw, vr, 30)
    assert torch.equal(i.cpu(), ri) and torch.equal(s.cpu(), rs) and torch.equal(w.cpu(), rw)
    lb, ab, vb = loc.bfloat16(), torch.softmax(torch.randn(N, Lq, 8, 16, generator=g), -1).view(N, Lq, 8, 4, 4).bfloat16(), vr.bfloat16()
    s, w, i = capi.topk_samples(lb.to(DEV), ab.to(DEV), vb.to(DEV), 30)
    ri = oracle.decoder_sample_topk(lb, ab, vb, 30)[2]
    assert torch.equal(i.cpu(), ri) and torch.equal(w.cpu(), torch.gather(ab.view(N, Lq, -1), 2, ri))
    ref_s = torch.gather((lb.float() / vb.float()[:, None, None, :, None, :]).bfloat16().view(N, Lq, -1, 2), 2,
                         ri.unsqueeze(-1).repeat(1, 1, 1, 2))
    assert torch.equal(s.cpu(), ref_s)


# ------------------------------------------------------------------------------------------------
# randomised shapes: fast and generic kernels, ragged levels, single rows/columns, Lq = 1
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(24))
def test_random_shapes_against_the_c_oracle(capi, oracle, seed):
    """Seeded random problem shapes (heads, channels, levels x points, level sizes down to 1x1, 1..7 frames,
    1..97 queries, locations up to 20 % outside the map) through the C ABI in fp32, against the fp64 C
    restatement of the reference kernels' arithmetic (oracle/msda_oracle.c) on the same fp32-rounded inputs."""
    rng = np.random.default_rng(1000 + seed)
    fast = seed % 2 == 0                       # D = 32 and L*P = 16 -> the fast kernels, else the generic ones
    if fast:
        M, D = int(rng.choice([4, 8, 12, 16])), 32
        L, P = [(4, 4), (2, 8), (1, 16), (8, 2), (16, 1)][int(rng.integers(0, 5))]
    else:
        M, D = int(rng.integers(1, 7)), int(rng.choice([1, 3, 8, 16, 33, 48, 64]))
        L, P = int(rng.integers(1, 6)), int(rng.integers(1, 6))
    levels = [(int(rng.integers(1, 14)), int(rng.integers(1, 14))) for _ in range(L)]
    N, Lq = int(rng.integers(1, 8)), int(rng.integers(1, 98))
    S = sum(h * w for h, w in levels)
    g = torch.Generator().manual_seed(seed)
    value = torch.randn(N, S, M, D, generator=g)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.4 - 0.2
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    go = torch.randn(N, Lq, M * D, generator=g)
    shapes, lsi = oracle.level_tensors(levels)
    lib = capi.load()
    sh = capi.make_shape(N, S, Lq, M, D, L, P, capi.MSDA_F32, capi.MSDA_F32)
    assert lib.msda_kernel_name(C.byref(sh), 0).startswith(b"fwd_fast" if fast else b"fwd_generic")
    out = capi.forward(value.to(DEV), shapes.to(DEV), lsi.to(DEV), loc.to(DEV), aw.to(DEV))
    gv, gl, ga = capi.backward(value.to(DEV), shapes.to(DEV), lsi.to(DEV), loc.to(DEV), aw.to(DEV), go.to(DEV))
    ref_out = torch.as_tensor(oracle.c_forward(value.double(), shapes, lsi, loc.double(), aw.double()))
    rv, rl, ra = (torch.as_tensor(x) for x in oracle.c_backward(value.double(), shapes, lsi, loc.double(), aw.double(), go.double()))
    assert oracle.max_rel(out.cpu(), ref_out.view_as(out.cpu())) < FWD_TOL
    check_grads(oracle, (gv, gl, ga), (rv, rl, ra), loc, levels, GRAD_TOL, stress=True)


# ------------------------------------------------------------------------------------------------
# fp16 (generic kernels) and 16-bit values at non-TCE geometry
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("vdt,ldt", [(torch.float16, torch.float32), (torch.float16, torch.float16),
                                     (torch.bfloat16, torch.float32)])
def test_sixteen_bit_generic_kernels(capi, oracle, vdt, ldt):
    """Half-precision values take the generic kernels (fp32 accumulation, fp32 scratch for grad_value):
    TCE head geometry in fp16, and a 48-channel geometry in bf16 that the fast kernels do not cover."""
    D = 32 if vdt == torch.float16 else 48
    levels = [(12, 20), (6, 10), (3, 5), (2, 3)]
    inp = oracle.make_inputs("model", 2, levels, M=4, D=D, P=4, seed=31)
    t = dev_inputs(inp, vdt, ldt)
    lib = capi.load()
    sh = capi.shape_of(t["value"], t["loc"], 4)
    assert lib.msda_kernel_name(C.byref(sh), 0) == b"fwd_generic"
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    gv, gl, ga = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(t["value"].cpu().double(), inp["shapes"], t["loc"].cpu().double(),
                                                      t["aw"].cpu().double(), t["grad_out"].cpu().double())
    tol = 2e-3 if vdt == torch.float16 else BF16_TOL          # fp16: 11-bit mantissa outputs
    assert oracle.max_rel(out.float().cpu(), ref_out) < tol
    eps = 1e-4 if ldt == torch.float32 else 5e-3
    check_grads(oracle, (gv.float(), gl.float(), ga.float()), (rv, rl, ra), t["loc"].float().cpu(), levels, tol if ldt == torch.float32 else 1e-2,
                eps_px=eps, stress=True)


# ------------------------------------------------------------------------------------------------
# non-finite / overflowing sampling locations
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", ["fast", "generic"])
def test_non_finite_and_overflowing_locations_contribute_nothing(capi, oracle, path):
    """A sampling point whose location is inf / NaN / overflows when scaled is gated off by the reference
    (ms_deform_im2col_cuda.cuh:441: all four comparisons are false): the output and the gradients of the
    other points must be unaffected and the gated point's own gradients are zero.  The fast kernels used to
    turn such a point into NaN through inf - inf in the fractional part."""
    levels = [(12, 20), (6, 10), (3, 5), (2, 3)]
    D = 32 if path == "fast" else 16                     # D != 32 -> generic kernels
    inp = oracle.make_inputs("model", 2, levels, D=D, seed=14)
    loc = inp["loc"].clone()
    bad = torch.tensor([float("inf"), float("-inf"), float("nan"), 1e38, -1e38, 3.0e30])
    flat = loc.view(-1, 2)
    idx = torch.arange(0, flat.shape[0], 17)
    flat[idx, 0] = bad[torch.arange(idx.numel()) % bad.numel()]
    flat[idx[::2], 1] = bad[(torch.arange(idx[::2].numel()) + 2) % bad.numel()]
    t = dev_inputs(dict(inp, loc=loc), torch.float32)
    lib = capi.load()
    sh = capi.shape_of(t["value"], t["loc"], len(levels))
    assert (b"fast" in lib.msda_kernel_name(C.byref(sh), 0)) == (path == "fast")
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
    assert float(torch.nan_to_num(gl.cpu()[bad_pt], nan=1.0).abs().max()) == 0.0
    ok = ~bad_pt
    err = (gl.cpu()[ok].double() - rl[ok]).abs().max() / rl.abs().max()
    assert float(err) < 1e-3          # boundary points excluded elsewhere; here only the blast radius matters


# ------------------------------------------------------------------------------------------------
# merging of the duplicate corners of one (query, head, level)  (msda_kernels.cuh: merge_level)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,loc_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32),
                                             (torch.bfloat16, torch.bfloat16)])
def test_overlapping_points_within_a_level(capi, oracle, dtype, loc_dtype):
    """The fast kernels merge the corners that several points of one (query, head, level) share before they load /
    reduce them.  Drive every overlap pattern on purpose: the four points of a level are placed on a small integer
    lattice around the reference point (offsets 0, +-1 px in x and y, in every order, so that blocks coincide, overlap
    by a column, by a row, by one diagonal corner, or not at all), including queries at the borders (corners dropped
    by zero padding), points outside the map (gated off) between two overlapping ones, and zero attention weights."""
    levels = [(12, 20), (6, 10), (5, 1), (2, 3)]           # a level of width 1: pixel offsets 1 and W alias as indices
    g = torch.Generator().manual_seed(31)
    N, M, P, L = 3, 8, 4, 4
    S = sum(h * w for h, w in levels)
    inp = oracle.make_inputs("model", N, levels, seed=31)
    Lq = inp["loc"].shape[1]
    ref = oracle.reference_points_grid(levels)                                   # [S, 2], pixel centres
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32)
    step = torch.randint(-1, 2, (N, Lq, M, L, P, 2), generator=g).float()        # lattice offsets in pixels
    frac = 0.25 + 0.5 * torch.rand(N, Lq, M, L, 1, 2, generator=g)               # same sub-pixel phase for the 4 points
    loc = ref[None, :, None, None, None, :] + (step + frac - 0.5) / wh[None, None, None, :, None, :]
    far = torch.rand(N, Lq, M, L, P, generator=g) < 0.1                          # some points far outside
    loc[far] = 3.0
    aw = inp["aw"].clone()
    aw[torch.rand(aw.shape, generator=g) < 0.1] = 0.0
    src = dict(inp, loc=loc.contiguous(), aw=aw.contiguous())
    t = dev_inputs(src, dtype, loc_dtype)
    lib = capi.load()
    sh = capi.shape_of(t["value"], t["loc"], L)
    assert b"fast" in lib.msda_kernel_name(C.byref(sh), 0)
    out = capi.forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])
    got = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"])
    ref_out, rv, rl, ra = oracle.core_pytorch_fwd_bwd(t["value"].cpu().double(), inp["shapes"], t["loc"].cpu().double(),
                                                      t["aw"].cpu().double(), t["grad_out"].cpu().double())
    f_tol, g_tol = (FWD_TOL, GRAD_TOL) if dtype == torch.float32 else (BF16_TOL, BF16_TOL)
    assert oracle.max_rel(out.float().cpu(), ref_out) < f_tol
    eps = 1e-4 if loc_dtype == torch.float32 else 2e-2
    check_grads(oracle, tuple(x.float() for x in got), (rv, rl, ra), t["loc"].float().cpu(), levels, g_tol, eps_px=eps, stress=True)
    if dtype == torch.bfloat16:      # the direct 16-bit reduction path (4 lanes x 8 channels, one lane owns a whole level)
        gv2, gl2, ga2 = capi.backward(t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"], t["grad_out"], use_workspace=False)
        assert oracle.max_rel(gv2.float().cpu(), rv) < 1.5e-1
        assert oracle.max_rel(ga2.float().cpu(), ra) < BF16_TOL
