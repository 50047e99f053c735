"""The oracle against the golden vectors produced by the REFERENCE's own
ms_deform_attn_core_pytorch (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden

NAMES = golden_names()


def test_golden_vectors_present():
    assert len(NAMES) >= 9, "tests/golden/*.npz missing — run oracle/gen_golden.py in the build container"


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_f64_matches_reference_golden(oracle, name):
    g = load_golden(name)
    out = oracle.c_forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], np.float64)
    assert oracle.max_rel(out, g["out"]) < 1e-12
    gv, gl, ga = oracle.c_backward(g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], g["grad_out"], np.float64)
    assert oracle.max_rel(gv, g["grad_value"]) < 1e-12
    assert oracle.max_rel(ga, g["grad_aw"]) < 1e-12
    # stress cases contain exact-integer pixel coordinates on purpose: grad_sampling_loc is
    # discontinuous there, so only non-boundary points are comparable
    err, l2, frac = oracle.grad_loc_errors(gl, g["grad_loc"], g["loc"], g["shapes"].tolist(), eps_px=1e-9)
    assert err < 1e-11, (err, l2, frac)
    if str(g["kind"]) != "stress":
        assert l2 < 1e-11, (err, l2, frac)


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_f32_within_fp32_tolerance(oracle, name):
    g = load_golden(name)
    out = oracle.c_forward(g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], np.float32)
    assert oracle.max_rel(out, g["out"]) < 1e-5
    gv, gl, ga = oracle.c_backward(g["value"], g["shapes"], g["lsi"], g["loc"], g["aw"], g["grad_out"], np.float32)
    assert oracle.max_rel(gv, g["grad_value"]) < 1e-4
    assert oracle.max_rel(ga, g["grad_aw"]) < 1e-4
    err, l2, _ = oracle.grad_loc_errors(gl, g["grad_loc"], g["loc"], g["shapes"].tolist(), eps_px=1e-3)
    assert err < 1e-4, err


@pytest.mark.parametrize("name", NAMES)
def test_pytorch_restatement_matches_reference_golden(oracle, name):
    g = load_golden(name)
    t = {k: torch.from_numpy(np.asarray(v)) for k, v in g.items() if k not in ("kind", "meta")}
    out, gv, gl, ga = oracle.core_pytorch_fwd_bwd(t["value"].double(), t["shapes"], t["loc"], t["aw"], t["grad_out"])
    assert oracle.max_rel(out, g["out"]) < 1e-12
    assert oracle.max_rel(gv, g["grad_value"]) < 1e-12
    assert oracle.max_rel(gl, g["grad_loc"]) < 1e-12
    assert oracle.max_rel(ga, g["grad_aw"]) < 1e-12


def test_reference_test_py_checks_on_oracle(oracle):
    """models/ops/test.py:31-60: forward equality double / float at its geometry."""
    inp = oracle.make_inputs("testpy", 1, oracle.LEVELS_TEST_PY, Lq=2, M=2, D=2, P=2, seed=3, dtype=torch.float64)
    ref = oracle.core_pytorch(inp["value"], inp["shapes"], inp["loc"], inp["aw"])
    out64 = oracle.c_forward(inp["value"], inp["shapes"], inp["lsi"], inp["loc"], inp["aw"], np.float64)
    assert torch.allclose(torch.from_numpy(out64), ref)                       # test.py:40 default tolerances
    out32 = oracle.c_forward(inp["value"], inp["shapes"], inp["lsi"], inp["loc"], inp["aw"], np.float32)
    assert torch.allclose(torch.from_numpy(out32).double(), ref, rtol=1e-2, atol=1e-3)   # test.py:56


def test_oracle_empty_and_all_outside(oracle):
    shapes, lsi = oracle.level_tensors([(3, 4)])
    value = torch.randn(2, 12, 2, 4, dtype=torch.float64)
    loc = torch.full((2, 3, 2, 1, 2, 2), 5.0, dtype=torch.float64)            # everything outside
    aw = torch.full((2, 3, 2, 1, 2), 0.5, dtype=torch.float64)
    out = oracle.c_forward(value, shapes, lsi, loc, aw)
    assert np.all(out == 0)
    gv, gl, ga = oracle.c_backward(value, shapes, lsi, loc, aw, np.ones_like(out))
    assert np.all(gv == 0) and np.all(gl == 0) and np.all(ga == 0)
    # Lq = 0
    out0 = oracle.c_forward(value, shapes, lsi, loc[:, :0], aw[:, :0])
    assert out0.shape == (2, 0, 8)


def test_oracle_3d_reduces_to_2d_at_frame_centres(oracle):
    """f = (k+0.5)/G puts the blend exactly on frame k: the 3-D operator must equal the 2-D operator
    reading frame k for every query."""
    levels = [(4, 5), (2, 3)]
    inp = oracle.make_inputs("stress", 3, levels, Lq=6, M=2, D=4, P=2, seed=11, dtype=torch.float64)
    N = 3
    for k in range(N):
        f = torch.full(inp["loc"].shape[:-1] + (1,), (k + 0.5) / N, dtype=torch.float64)
        loc3 = torch.cat([inp["loc"], f], -1)
        out3 = oracle.c_forward(inp["value"], inp["shapes"], inp["lsi"], loc3, inp["aw"])
        vk = inp["value"][k:k + 1].expand(N, -1, -1, -1).contiguous()
        out2 = oracle.c_forward(vk, inp["shapes"], inp["lsi"], inp["loc"], inp["aw"])
        assert oracle.max_rel(out3, out2) < 1e-13


def test_oracle_3d_backward_against_finite_differences(oracle):
    levels = [(4, 5), (2, 3)]
    N = 3
    inp = oracle.make_inputs("testpy", N, levels, Lq=3, M=2, D=3, P=2, seed=5, dtype=torch.float64, three_d=True)
    loc = inp["loc"].clone()
    loc[..., 2] = loc[..., 2].clamp(0.05, 0.95)
    # keep every coordinate away from cell boundaries so finite differences are valid
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float64)
    px = loc[..., :2] * wh[None, None, None, :, None, :] - 0.5
    frac = px - px.floor()
    px = px.floor() + frac.clamp(0.1, 0.9)
    loc[..., :2] = (px + 0.5) / wh[None, None, None, :, None, :]
    fi = loc[..., 2] * N - 0.5
    loc[..., 2] = (fi.floor() + (fi - fi.floor()).clamp(0.1, 0.9) + 0.5) / N
    go = inp["grad_out"].numpy()

    def f(v, l, a):
        return float((oracle.c_forward(v, inp["shapes"], inp["lsi"], l, a) * go).sum())

    gv, gl, ga = oracle.c_backward(inp["value"], inp["shapes"], inp["lsi"], loc, inp["aw"], go)
    rng = np.random.default_rng(0)
    eps = 1e-6
    for _ in range(12):
        idx = tuple(rng.integers(0, s) for s in loc.shape)
        lp, lm = loc.clone(), loc.clone()
        lp[idx] += eps; lm[idx] -= eps
        num = (f(inp["value"], lp, inp["aw"]) - f(inp["value"], lm, inp["aw"])) / (2 * eps)
        ana = gl[idx]
        if idx[-1] == 2:
            ana = ana * N / (N - 1)      # the fork scales d/df by (G-1), the forward by G (.cuh:241)
        assert abs(num - ana) < 1e-5 * max(1.0, abs(num)), (idx, num, ana)
    for _ in range(6):
        idx = tuple(rng.integers(0, s) for s in inp["aw"].shape)
        ap, am = inp["aw"].clone(), inp["aw"].clone()
        ap[idx] += eps; am[idx] -= eps
        num = (f(inp["value"], loc, ap) - f(inp["value"], loc, am)) / (2 * eps)
        assert abs(num - ga[idx]) < 1e-6 * max(1.0, abs(num))
        idx = tuple(rng.integers(0, s) for s in inp["value"].shape)
        vp, vm = inp["value"].clone(), inp["value"].clone()
        vp[idx] += eps; vm[idx] -= eps
        num = (f(vp, loc, inp["aw"]) - f(vm, loc, inp["aw"])) / (2 * eps)
        assert abs(num - gv[idx]) < 1e-6 * max(1.0, abs(num))


# ------------------------------------------------------------------------------------------------
# rows f2 / f3: the restatements of the passes either side of the op
# ------------------------------------------------------------------------------------------------
def test_decoder_sample_topk_restatement_matches_the_reference_calls(oracle):
    """models/tce_deformable_transformer.py:752-758 with its literal torch.topk vs the stable-sort
    restatement the CUDA kernel is checked against: identical wherever weights are distinct (softmax of
    random logits), and for all-equal weights (the model at initialisation: attention_weights.weight = 0,
    ms_deform_attn.py:72-73) the restatement picks the first k candidates."""
    g = torch.Generator().manual_seed(0)
    N, Lq, M, L, P = 3, 7, 8, 4, 4
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    vr = 0.5 + 0.5 * torch.rand(N, L, 2, generator=g)
    s1, w1, i1 = oracle.decoder_sample_topk(loc, aw, vr, 30)
    s2, w2, i2 = oracle.decoder_sample_topk_reference_calls(loc, aw, vr, 30)
    assert torch.equal(w1, w2) and torch.equal(i1, i2) and torch.equal(s1, s2)
    assert s1.shape == (N, Lq, 30, 2) and w1.shape == (N, Lq, 30)
    assert bool((w1[..., :-1] >= w1[..., 1:]).all())
    flat = torch.full((N, Lq, M, L, P), 1.0 / 16)
    s3, w3, i3 = oracle.decoder_sample_topk(loc, flat, vr, 30)
    assert torch.equal(i3, torch.arange(30).expand(N, Lq, 30))
    assert torch.equal(s3, (loc / vr[:, None, None, :, None, :]).reshape(N, Lq, -1, 2)[:, :, :30])


def test_value_epilogue_restatement(oracle):
    g = torch.Generator().manual_seed(1)
    v = torch.randn(2, 11, 256, generator=g)
    m = torch.rand(2, 11, generator=g) < 0.3
    out = oracle.value_epilogue(v, m)
    assert torch.equal(out[m], torch.zeros_like(out[m])) and torch.equal(out[~m], v[~m])
    assert oracle.value_epilogue(v, m, torch.bfloat16).dtype == torch.bfloat16
    assert torch.equal(oracle.value_epilogue(v), v)


def test_baseline_config0_cpu_forward(oracle):
    """BASELINE.json configs[0], the reference's own CPU-runnable case: standalone forward through
    ms_deform_attn_core_pytorch, B*T = 5 frames, levels 45x80 / 23x40 / 12x20 / 6x10, 8 heads x 32 ch,
    4 points, fp32 synthetic tensors.  The PyTorch restatement in fp32 agrees with the fp64 C restatement
    of the CUDA kernels' arithmetic to fp32 rounding (1e-6)."""
    inp = oracle.make_inputs("model", 5, oracle.LEVELS_A, seed=0)
    out32 = oracle.core_pytorch(inp["value"], inp["shapes"], inp["loc"], inp["aw"])
    assert out32.shape == (5, 4820, 256) and out32.dtype == torch.float32
    ref = torch.as_tensor(oracle.c_forward(inp["value"].double(), inp["shapes"], inp["lsi"], inp["loc"].double(),
                                           inp["aw"].double())).view_as(out32)
    assert oracle.max_rel(out32, ref) < 1e-5
