"""Seeded synthetic workloads of SURVEY.md 8(d) — input generators only, no arithmetic of the operator.

Shared by the benchmark (bench.py, tools/kbench.py ...) and by the tests; kept OUT of oracle/ so that the
measured paths never import the checker.  Level sets: A 360x640, B 384x640 (A after pad-to-32), C 512x896
(SURVEY.md 8), and the 2-level set of the reference's own test (models/ops/test.py:23).
"""
from __future__ import annotations

import numpy as np
import torch

# level sets of SURVEY.md section 8
LEVELS_A = [(45, 80), (23, 40), (12, 20), (6, 10)]       # 360x640  -> S = 4820
LEVELS_B = [(48, 80), (24, 40), (12, 20), (6, 10)]       # 384x640  -> S = 5100
LEVELS_C = [(64, 112), (32, 56), (16, 28), (8, 14)]      # 512x896  -> S = 9520
LEVELS_TEST_PY = [(6, 4), (3, 2)]                        # models/ops/test.py:23


def level_tensors(levels, device="cpu"):
    shapes = torch.as_tensor(levels, dtype=torch.long, device=device)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return shapes, lsi


def reference_points_grid(levels, dtype=torch.float32):
    """Pixel centres of every level, normalised — get_reference_points,
    models/tce_deformable_transformer.py:572-589 with valid_ratios = 1.  -> [S, 2] (x, y)."""
    pts = []
    for (h, w) in levels:
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, dtype=dtype),
                                torch.linspace(0.5, w - 0.5, w, dtype=dtype), indexing="ij")
        pts.append(torch.stack((xs.reshape(-1) / w, ys.reshape(-1) / h), -1))
    return torch.cat(pts, 0)


def make_inputs(kind, N, levels, Lq=None, M=8, D=32, P=4, seed=0, dtype=torch.float32, three_d=False):
    """kind: 'testpy' (models/ops/test.py:33-36 — uniform locations, worst-case locality),
    'model' (encoder-like: reference point = the query's own pixel centre, offsets = the module's
    init pattern k*dir_m/(W_l,H_l) + N(0, 0.5 px) jitter, softmax weights; primary for roofline),
    'stress' (U[-0.1,1.1) locations incl. exact-integer and all-outside points).
    Returns dict(value, shapes, lsi, loc, aw, grad_out) on CPU."""
    g = torch.Generator().manual_seed(seed)
    L = len(levels)
    S = sum(h * w for h, w in levels)
    shapes, lsi = level_tensors(levels)
    LD = 3 if three_d else 2
    if kind == "model":
        if Lq is None:
            Lq = S
        value = torch.randn(N, S, M, D, generator=g)
        ref = reference_points_grid(levels)                       # [S, 2]
        if Lq != S:
            idx = torch.randint(0, S, (Lq,), generator=g)
            ref = ref[idx]
        thetas = torch.arange(M, dtype=torch.float32) * (2.0 * np.pi / M)
        dirs = torch.stack([thetas.cos(), thetas.sin()], -1)
        dirs = dirs / dirs.abs().max(-1, keepdim=True)[0]         # [M, 2]  (ms_deform_attn.py:63-68)
        k = torch.arange(1, P + 1, dtype=torch.float32)           # point p -> (p+1) px
        offs = dirs[:, None, None, :] * k[None, None, :, None]    # [M,1,P,2]
        offs = offs.expand(M, L, P, 2) + 0.5 * torch.randn(N, Lq, M, L, P, 2, generator=g)
        wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32)   # [L,2] (W,H)
        loc = ref[None, :, None, None, None, :] + offs / wh[None, None, None, :, None, :]
        aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    elif kind == "testpy":
        if Lq is None:
            Lq = S
        value = torch.rand(N, S, M, D, generator=g) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
        aw = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
        aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    elif kind == "stress":
        if Lq is None:
            Lq = 64
        value = torch.randn(N, S, M, D, generator=g)
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.2 - 0.1
        wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32)
        flat = loc.view(N, Lq, M, L, P, 2)
        # a slice of exact pixel-centre / pixel-edge coordinates and of far-outside points
        flat[:, 0::7, :, :, 0, :] = (torch.floor(flat[:, 0::7, :, :, 0, :] * wh[None, None, None]) + 0.5) / wh[None, None, None]
        flat[:, 1::7, :, :, 1 % P, :] = torch.floor(flat[:, 1::7, :, :, 1 % P, :] * wh[None, None, None]) / wh[None, None, None]
        flat[:, 2::7, :, :, 2 % P, :] = 3.0
        flat[:, 3::7, :, :, 3 % P, :] = -2.0
        aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    else:
        raise ValueError(kind)
    if three_d:
        f = torch.rand(N, Lq, M, L, P, 1, generator=g) * 1.2 - 0.1
        loc = torch.cat([loc, f], -1)
    grad_out = torch.randn(N, Lq, M * D, generator=g)
    return dict(value=value.to(dtype).contiguous(), shapes=shapes, lsi=lsi, loc=loc.to(dtype).contiguous(),
                aw=aw.to(dtype).contiguous(), grad_out=grad_out.to(dtype).contiguous())
