"""Module-level timing: MSDeformAttn forward + backward at the encoder shape of BASELINE.json configs[1]
(N = 10 frames, S = Lq = 4820, d_model 256, 10 % padded pixels) for the reference composition
(models/ops/modules/ms_deform_attn.py:95-115: torch masked_fill, softmax, location arithmetic around the op)
and for this repo's fused options (rows f1 + f2).  The four nn.Linear layers (cuBLAS) are inside the timed
region in every variant.  CUDA events, 3 rotating input sets.  Developer tool.   python tools/modbench.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("tce-rvos_b200")
from tools import synthetic  # noqa: E402


def main():
    dev = "cuda:0"
    levels = synthetic.LEVELS_A
    N, S = 10, sum(h * w for h, w in levels)
    shapes, lsi = synthetic.level_tensors(levels, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    torch.manual_seed(0)
    mod = pkg.MSDeformAttn().to(dev)
    with torch.no_grad():                      # a trained-like module: non-trivial offsets and weights
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.1)
    ref_pts = synthetic.reference_points_grid(levels).to(dev)[None, :, None, :].expand(N, S, 4, 2).contiguous()
    sets = [dict(q=torch.randn(N, S, 256, device=dev, generator=g, requires_grad=True),
                 src=torch.randn(N, S, 256, device=dev, generator=g, requires_grad=True),
                 pm=torch.rand(N, S, device=dev, generator=g) < 0.1,
                 go=torch.randn(N, S, 256, device=dev, generator=g)) for _ in range(3)]

    def step(i):
        s = sets[i % 3]
        out = mod(s["q"], ref_pts, s["src"], shapes, lsi, s["pm"])[0]
        out.backward(s["go"])
        s["q"].grad = s["src"].grad = None
        mod.zero_grad(set_to_none=True)

    def timed(reps=20):
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i, (a, b) in enumerate(evs):
            a.record(); step(i); b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
        return ts[len(ts) // 2]

    variants = [("reference composition (torch masked_fill / softmax / locations + op)", {}),
                ("fused prologue (f1), locations+weights materialised (decoder use)", dict(fused_prologue=True, return_sampling=True)),
                ("fused prologue (f1), not materialised (encoder / FTF use)", dict(fused_prologue=True, return_sampling=False)),
                ("f1 + fused value epilogue (f2), fp32 value", dict(fused_prologue=True, return_sampling=False, fused_value_epilogue=True)),
                ("f1 + f2 with bf16 value", dict(fused_prologue=True, return_sampling=False, fused_value_epilogue=True, value_dtype=torch.bfloat16))]
    base = None
    for name, opts in variants:
        mod.fused_prologue, mod.return_sampling, mod.fused_value_epilogue, mod.value_dtype = False, True, False, None
        for k, v in opts.items():
            setattr(mod, k, v)
        us = timed()
        base = base or us
        print(f"{name:75s} {us:8.1f} us fwd+bwd   {base / us:4.2f}x")


if __name__ == "__main__":
    main()
