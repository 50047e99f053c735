"""Decoder / FTF sized calls (36 frames, Lq = 5 and 8): what one call costs when it is launched
eagerly from Python and what the same call costs as a node of a CUDA graph (a decoder layer or a whole forward
captured once), i.e. how much of the small calls is launch overhead and how much is kernel.

    python tools/smallcall_bench.py [--frames 36] [--calls 12]
"""
import argparse
import importlib
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("tce-rvos_b200")
from tools import synthetic as O  # noqa: E402


def timed(fn, reps):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) * 1e3 for a, b in evs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=36)
    ap.add_argument("--calls", type=int, default=12, help="op calls per graph (12 = the op calls of one 4+4-layer forward)")
    a = ap.parse_args()
    dev = "cuda:0"
    ext = pkg.load_extension()
    rows = []
    for Lq in (5, 8):
        inp = O.make_inputs("model", a.frames, O.LEVELS_A, Lq=Lq, seed=Lq)
        t = {k: inp[k].to(dev) for k in ("value", "loc", "aw", "grad_out", "shapes", "lsi")}
        args = (t["value"], t["shapes"], t["lsi"], t["loc"], t["aw"])

        def fwd():
            return ext.ms_deform_attn_forward(*args, 64, False)

        def bwd():
            return ext.ms_deform_attn_backward(*args, t["grad_out"], 64, False)

        for name, f in (("forward", fwd), ("backward", bwd)):
            for _ in range(5):
                f()
            torch.cuda.synchronize()
            eager = timed(f, 50)                       # one call per event pair: includes the launch gap
            eager_b2b = timed(lambda: [f() for _ in range(a.calls)], 20) / a.calls      # back to back from Python
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                f()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep = [f() for _ in range(a.calls)]
            g.replay()
            torch.cuda.synchronize()
            graphed = timed(g.replay, 20) / a.calls
            rows.append({"Lq": Lq, "pass": name, "eager_us_per_call": eager, "eager_back_to_back_us_per_call": eager_b2b,
                         "in_cuda_graph_us_per_call": graphed})
            print(f"Lq={Lq} {name:8s}: eager {eager:6.1f} us/call, {a.calls} back-to-back eager calls {eager_b2b:6.1f} us/call, "
                  f"as nodes of one CUDA graph {graphed:6.1f} us/call", flush=True)
    print(json.dumps({"frames": a.frames, "calls_per_graph": a.calls, "rows": rows}))


if __name__ == "__main__":
    main()
