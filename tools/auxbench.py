"""Timing of the two small passes either side of the op (rows f2 / f3) against the torch calls the
reference makes (models/ops/modules/ms_deform_attn.py:96-97; models/tce_deformable_transformer.py:752-758).
CUDA events, L2 flushed between repetitions by rotating inputs (f2) — f3 is latency-bound and tiny.
Developer tool.   python tools/auxbench.py"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("tce-rvos_b200")
capi = importlib.import_module("tce-rvos_b200.capi")


def timed(f, reps=20, warm=3, per_graph=8):
    """Median time of one call.  The calls are replayed from a CUDA graph holding `per_graph` of them so
    that the GPU, not the Python launch path, is what is timed (these kernels run for 10-50 us).
    AUXBENCH_NO_GRAPH=1 (ncu captures): plain launches, 2 calls."""
    for i in range(warm):
        f(i)
    torch.cuda.synchronize()
    if os.environ.get("AUXBENCH_NO_GRAPH"):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(0); f(1); b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / 2
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            keep = [f(i) for i in range(per_graph)]
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); graph.replay(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 / per_graph for a, b in evs)
    del keep
    return ts[len(ts) // 2]


def main():
    dev = "cuda:0"
    peak = 6550.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    g = torch.Generator(device=dev).manual_seed(0)
    # ---- f2 at config 2: value [10, 4820, 256]; 4 rotating sets (4 x 49 MB + outputs > L2) ----
    N, S, C = 10, 4820, 256
    sets = [torch.randn(N, S, C, device=dev, generator=g) for _ in range(4)]
    masks = [(torch.rand(N, S, device=dev, generator=g) < 0.1) for _ in range(4)]
    nbytes = N * S * C * 4
    for name, f, moved in [
        ("torch masked_fill (reference, out of place)     ", lambda i: sets[i % 4].masked_fill(masks[i % 4][..., None], 0.0), 2 * nbytes),
        ("value_epilogue fp32 -> fp32                      ", lambda i: capi.value_epilogue(sets[i % 4], masks[i % 4]), 2 * nbytes),
        ("torch masked_fill + .to(bfloat16)                ", lambda i: sets[i % 4].masked_fill(masks[i % 4][..., None], 0.0).to(torch.bfloat16), 1.5 * nbytes),
        ("value_epilogue fp32 -> bf16                      ", lambda i: capi.value_epilogue(sets[i % 4], masks[i % 4], torch.bfloat16), 1.5 * nbytes),
    ]:
        us = timed(f)
        print(f"f2  {name} {us:8.1f} us   {moved / us / 1e3:7.0f} GB/s algorithmic = {moved / us / 1e3 / peak * 100:5.1f}% of {peak:.0f}")
    # ---- f3 at the decoder shapes ----
    for N, Lq in ((36, 5), (10, 5), (64, 300)):
        loc = torch.rand(N, Lq, 8, 4, 4, 2, device=dev, generator=g)
        aw = torch.softmax(torch.randn(N, Lq, 8, 16, device=dev, generator=g), -1).view(N, Lq, 8, 4, 4)
        vr = 0.5 + 0.5 * torch.rand(N, 4, 2, device=dev, generator=g)

        def ref(i):
            sl = loc / vr[:, None, None, :, None, :]
            wf = aw.view(N, Lq, -1)
            sf = sl.view(N, Lq, -1, 2)
            tw, ti = wf.topk(30, dim=2)
            return torch.gather(wf, 2, ti), torch.gather(sf, 2, ti.unsqueeze(-1).repeat(1, 1, 1, 2))

        us_ref = timed(ref)
        us = timed(lambda i: pkg.decoder_sample_topk(loc, aw, vr, 30))
        print(f"f3  N={N:3d} Lq={Lq:3d}: reference torch chain {us_ref:7.1f} us   decoder_sample_topk {us:7.1f} us   ({us_ref / us:.1f}x)")


if __name__ == "__main__":
    main()
