"""Kernel-level timing of the operator's passes (CUDA events on the launching stream, rotating
input sets, L2 flushed by data volume).  Developer tool; bench.py is the contract benchmark.

    python tools/kbench.py [--frames 10] [--levels A|B|C] [--reps 20] [--kinds model,testpy] [--dtypes f32,bf16]
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("tce-rvos_b200")
capi = importlib.import_module("tce-rvos_b200.capi")
from tools import synthetic as O  # noqa: E402  (seeded workload generator)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--levels", default="A")
    ap.add_argument("--lq", type=int, default=None)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--kinds", default="model,testpy")
    ap.add_argument("--dtypes", default="f32,bf16")
    ap.add_argument("--sets", type=int, default=3)
    ap.add_argument("--reference", action="store_true",
                    help="also time the reference's own CUDA extension (baseline/_ref/MSDA_reference_cuda.so, "
                         "built by tools/build_reference_cuda.py; fp32 only)")
    args = ap.parse_args()
    ref_ext = None
    if args.reference:
        import importlib.util
        path = os.path.join(ROOT, "baseline", "_ref", "MSDA_reference_cuda.so")
        spec = importlib.util.spec_from_file_location("MSDA_reference_cuda", path)
        ref_ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_ext)
    levels = {"A": O.LEVELS_A, "B": O.LEVELS_B, "C": O.LEVELS_C}[args.levels]
    dev = "cuda:0"
    ext = pkg.load_extension()
    lib = capi.load()
    peak = 6550.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for kind in args.kinds.split(","):
        base = [O.make_inputs(kind, args.frames, levels, Lq=args.lq, seed=s) for s in range(args.sets)]
        for dt in args.dtypes.split(","):
            tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[dt]
            sets = [dict(v=b["value"].to(dev, tdt), l=b["loc"].to(dev), a=b["aw"].to(dev), g=b["grad_out"].to(dev, tdt)) for b in base]
            shapes, lsi = base[0]["shapes"].to(dev), base[0]["lsi"].to(dev)
            N, S = base[0]["value"].shape[:2]
            Lq = base[0]["loc"].shape[1]
            sh = capi.make_shape(N, S, Lq, 8, 32, len(levels), 4, capi.dtype_code(tdt), capi.MSDA_F32)
            bf, bb = lib.msda_algorithmic_bytes(sh, 0), lib.msda_algorithmic_bytes(sh, 1)

            def fwd(s):
                return ext.ms_deform_attn_forward(s["v"], shapes, lsi, s["l"], s["a"], 64, False)

            def bwd(s):
                return ext.ms_deform_attn_backward(s["v"], shapes, lsi, s["l"], s["a"], s["g"], 64, False)

            passes = [("fwd", fwd, bf), ("bwd", bwd, bb)]
            if ref_ext is not None and dt == "f32":
                passes += [("ref-fwd", lambda s: ref_ext.ms_deform_attn_forward(s["v"], shapes, lsi, s["l"], s["a"], 64, False), bf),
                           ("ref-bwd", lambda s: ref_ext.ms_deform_attn_backward(s["v"], shapes, lsi, s["l"], s["a"], s["g"], 64, False), bb)]
            for name, f, nbytes in passes:
                for i in range(3):
                    f(sets[i % len(sets)])
                torch.cuda.synchronize()
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
                for i, (a, b) in enumerate(evs):
                    a.record(); f(sets[i % len(sets)]); b.record()
                torch.cuda.synchronize()
                ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
                med = ts[len(ts) // 2]
                qh = N * Lq * 8
                print(f"{kind:7s} {dt:5s} {name:7s}: median {med:8.1f} us  min {ts[0]:8.1f} us  "
                      f"{qh / med / 1e3:6.2f} Gqh/s  {nbytes / med / 1e3:7.0f} GB/s algorithmic = {nbytes / med / 1e3 / peak * 100:5.1f}% of {peak:.0f}",
                      flush=True)


if __name__ == "__main__":
    main()
