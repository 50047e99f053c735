#!/usr/bin/env python
"""bench.py — throughput of the MSDeformAttn hot path on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f32|bf16]
                    [--workload config2|sweep|model3|model4] [--frames F] [--no-graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (`config.workload` names the one that ran):

* config2 (default, the headline) — BASELINE.json configs[1]: one step = MSDeformAttn forward + backward over
  one batch of synthetic encoder-shaped input at TCE-RVOS's training shape: B*T = 2x5 = 10 frames per GPU,
  levels 45x80 / 23x40 / 12x20 / 6x10 (S = 4820), Lq = S, 8 heads x 32 channels, 4 points, fp32, through the
  reference-facing MSDeformAttnFunction.  Metric: query-heads (N*Lq*M) per second.
* sweep — BASELINE.json configs[4]: the same step at 512x896 (levels 64x112 ... 8x14, S = 9520) with
  --frames 64..1024 frames per GPU, sharded by frame over the ranks.  1024 frames at this size is past the
  reference's int32 offsets and its N % 64 chunking (models/ops/src/cuda/ms_deform_attn_cuda.cu:54-56,62,84).
* model3 / model4 — BASELINE.json configs[2] / configs[3]: the UNMODIFIED reference model (staged under
  baseline/_ref/tce_rvos_src by tools/model_dropin.py --stage) running on this repo's operator: 36-frame
  inference clip / Video-Swin-T training step under DDP (main.py:67-72).  Metric: frames per second.
  Prints an `unavailable` line and exits 0 when the staged tree is absent.

Keys of the line (config2 / sweep):

* `value`      inputs resident in HBM; the timed region replays CUDA graphs of the whole step (forward through
               MSDeformAttnFunction.apply + autograd backward, one graph per rotating input set), bracketed by
               barrier + synchronize, CUDA events on the launching stream, max over ranks.  `--no-graph` times
               the eager step instead.
* `roofline`   the dominant pass (backward: zero-fill + scatter kernel): algorithmic bytes / its CUDA-event time
               from the instrumented eager loop of the same steps that follows the timed region (events cannot be
               read back from inside a replayed graph), against MEASURED_PEAKS.json's copy bandwidth.
* `e2e`        same step through the C-ABI host-buffer call (msda_forward_backward_host): pinned HOST inputs,
               host<->device copies inside the timed region.
* `shard_parity` after the timed region every rank checks one frame of its own shard (forward and all three
               gradients) against the oracle in fp64 — the checker, outside every timed region.
* `cpu_baseline` / `--impl reference`: the reference's own ms_deform_attn_core_pytorch + autograd
               (models/ops/functions/ms_deform_attn_func.py:67-87; the staged file when present = kind
               "reference", else the oracle's restatement = kind "port") on a FIXED sample — one frame, all
               queries, min(16, allowed cores) host threads — so that the figure is reproducible across boxes.

Frames shard across ranks with no collective (weak scaling: every rank owns `frames_per_gpu` frames).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LEVELS_A = [(45, 80), (23, 40), (12, 20), (6, 10)]        # 360x640
LEVELS_C = [(64, 112), (32, 56), (16, 28), (8, 14)]       # 512x896
M, D, L, P = 8, 32, 4, 4
METRIC = "msdeformattn_fwd_bwd_query_heads_per_s"
UNIT = "Gquery-heads/s"
STAGED = os.path.join(ROOT, "baseline", "_ref", "tce_rvos_src")


class Workload:
    def __init__(self, name, frames):
        self.name = name
        if name == "config2":
            self.levels, self.frames, self.n_sets, self.base_frames = LEVELS_A, 10, 3, 10
            self.title = ("BASELINE.json configs[1]: MSDeformAttn fwd+bwd, B*T=2x5 frames/GPU, levels 45x80/23x40/12x20/6x10 "
                          "(S=4820), Lq=S, 8 heads x 32 ch, 4 levels x 4 points, model-like sampling locations (SURVEY.md 8d-2)")
        else:
            self.levels, self.frames, self.n_sets = LEVELS_C, int(frames or 64), 1
            self.base_frames = min(self.frames, 8)
            self.title = (f"BASELINE.json configs[4]: MSDeformAttn fwd+bwd scaling sweep, {self.frames} frames/GPU at 512x896, "
                          "levels 64x112/32x56/16x28/8x14 (S=9520), Lq=S, 8 heads x 32 ch, 4 levels x 4 points, model-like "
                          "sampling locations")
        self.S = sum(h * w for h, w in self.levels)

    def config(self, dtype, n_gpus):
        e = 4 if dtype == "f32" else 2         # value / out / grad_out / grad_value; locations and weights stay fp32
        step_mb = self.frames * self.S * (e * (768 + 512) + 4 * 1152) / 1e6      # SURVEY.md 8(d), Lq = S
        if self.n_sets > 1:
            l2 = f"{self.n_sets} rotating input sets, each {step_mb:.0f} MB per fwd+bwd > 126 MB L2"
        else:
            l2 = f"one input set of {step_mb:.0f} MB per fwd+bwd, far larger than the 126 MB L2"
        cfg = {"workload": self.title, "frames_per_gpu": self.frames, "S": self.S, "Lq": self.S, "heads": M, "channels": D,
               "levels": L, "points": P, "value_dtype": dtype, "loc_dtype": "f32",
               "sharding": f"frames x {n_gpus} ranks, no collective", "l2_policy": l2}
        if self.base_frames != self.frames:
            cfg["synthetic_frames"] = (f"{self.base_frames} distinct synthetic frames repeated on the device to {self.frames} "
                                       "(distinct memory, same statistics)")
        return cfg


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the benchmark runs."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index, period=0.02):
        self.index, self.period = index, period
        self.samples, self.max_mhz = [], None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:                                    # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util, r))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()

    def summary(self):
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "note": "NVML unavailable: " + getattr(self, "err", "no samples")}
        mhz = [s[0] for s in self.samples]
        reasons = set()
        for _, _, r in self.samples:
            for name, bit in {**self.BAD, **self.NOTE}.items():
                if r & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(mhz), "sm_min_mhz": min(mhz), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(mhz),
                "window": "warm-up + timed region + post-roll of the same step (NVML, 20 ms period)"}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU implementation of the path
# ------------------------------------------------------------------------------------------------
def reference_cpu_function():
    """(fn, kind).  The reference's unmodified ms_deform_attn_core_pytorch when its python tree is staged under
    baseline/_ref (models/ops/functions/ms_deform_attn_func.py:67-87; its module-level `import
    MultiScaleDeformableAttention_update` is satisfied by an empty stub so that no extension of this repo is loaded on
    this arm), else the oracle's restatement of the same function."""
    path = os.path.join(STAGED, "models", "ops", "functions", "ms_deform_attn_func.py")
    if os.path.exists(path):
        try:
            import importlib.util
            name = "MultiScaleDeformableAttention_update"
            had = sys.modules.get(name)
            sys.modules[name] = types.ModuleType(name)
            try:
                spec = importlib.util.spec_from_file_location("_tce_rvos_ref_ms_deform_attn_func", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
            finally:
                if had is not None:
                    sys.modules[name] = had
                else:
                    del sys.modules[name]
            return mod.ms_deform_attn_core_pytorch, "reference"
        except Exception:
            pass
    from oracle import msda_oracle as O
    return O.core_pytorch, "port"


CPU_ARM_THREADS = 16      # fixed, so that the CPU figure (and the driver's GPU/CPU ratio) is the same on every box of the pool


def host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm always uses min(16, cores the process is allowed) threads: the
    1-GPU boxes of this pool expose 16 cores and the 8-GPU boxes 32, and grid_sample's backward gains only ~10 % from the
    second 16, so a fixed count keeps the figure comparable between the two (verdict r1: the ratio swung 2x between
    BENCH and SCALE with the sample and the thread count both floating)."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    n = max(1, min(n, CPU_ARM_THREADS))
    if torch.get_num_threads() != n:
        torch.set_num_threads(n)
    return torch.get_num_threads()


def time_cpu(wl, warmup, steps):
    """The FIXED sample of the CPU arm: one frame, all Lq = S queries of the workload, forward + autograd backward."""
    import torch
    from tools import synthetic
    threads = host_threads()
    fn, kind = reference_cpu_function()
    inp = synthetic.make_inputs("model", 1, wl.levels, M=M, D=D, P=P, seed=0, dtype=torch.float32)

    def step():
        v, l, a = (inp[k].clone().requires_grad_(True) for k in ("value", "loc", "aw"))
        out = fn(v, inp["shapes"], l, a)
        out.backward(inp["grad_out"])
        return out

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    qh = inp["loc"].shape[1] * M
    sample = (f"1 frame x all {inp['loc'].shape[1]} queries of the workload (same S, heads, points) per step, forward + autograd "
              f"backward of {'the reference file models/ops/functions/ms_deform_attn_func.py (staged, unmodified)' if kind == 'reference' else 'the oracle port of ms_deform_attn_core_pytorch'}, "
              f"torch {torch.__version__} CPU, {threads} threads, os.cpu_count()={os.cpu_count()}, {warmup} warm-up + {steps} timed steps")
    return qh * steps / dt / 1e9, dt / steps, threads, kind, sample


def run_reference(args, wl, rank):
    """--impl reference: rank 0 only; the other ranks exit 0 without work."""
    if rank != 0:
        return
    value, sec, threads, kind, sample = time_cpu(wl, max(1, args.warmup), args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl.config("f32", 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def count_line_updates(loc, levels):
    """(issued, unmerged): the number of 128-byte grad_value line updates one backward issues and the number the
    operator's definition asks for.  Unmerged: one per (query, head, point, corner) whose corner lies inside its level
    (the reference's gating, ms_deform_im2col_cuda.cuh:441 and :57-79).  Issued: the backward kernel folds a corner
    that falls on a pixel the PREVIOUS point of the same (query, head, level) also touches into that point's reduction
    (msda_kernels.cuh merge_level), so those are not sent."""
    import torch
    total = dup = 0
    for l, (H, W) in enumerate(levels):
        x = loc[:, :, :, l, :, 0] * W - 0.5
        y = loc[:, :, :, l, :, 1] * H - 0.5
        inside = (x > -1) & (y > -1) & (x < W) & (y < H)
        x0, y0 = torch.floor(x).long(), torch.floor(y).long()
        ids = []
        for dy in (0, 1):
            for dx in (0, 1):
                ok = inside & (x0 + dx >= 0) & (x0 + dx <= W - 1) & (y0 + dy >= 0) & (y0 + dy <= H - 1)
                ids.append(torch.where(ok, (y0 + dy) * W + (x0 + dx), torch.full_like(x0, -1)))
        ids = torch.stack(ids, -1)                                        # [N, Lq, M, P, 4] pixel index or -1
        total += int((ids >= 0).sum())
        cur, prev = ids[:, :, :, 1:, :, None], ids[:, :, :, :-1, None, :]
        dup += int(((cur >= 0) & (cur == prev)).any(-1).sum())
    return total - dup, total


# measured on this pool (profiles/r01_microbench_bulk_red.txt, gpurun_out/red_scaling.log): the chip sustains 47.5-49.4 G
# updates/s of 128-byte lines through red.global.add.v4.f32 or TMA bulk reductions, from 74 SMs upwards = ~6.1 TB/s
L2_REDUCTION_PEAK_GLINES = 47.5


def bind_to_gpu_numa_node(index):
    """Pin this process (and the first touch of its pinned buffers) to the CPUs NVML reports as local to the GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def shard_parity(wl, inp_cpu, dev_set, shapes, lsi, fn, tdt):
    """One frame of this rank's shard, forward + gradients on the device against the oracle in fp64 (the checker;
    outside every timed region).  Returns the four scale-relative errors and the tolerances they are held to."""
    import torch
    from oracle import msda_oracle as O
    f = 0
    v = dev_set["value"][f:f + 1].detach().clone().requires_grad_(True)
    l = dev_set["loc"][f:f + 1].detach().clone().requires_grad_(True)
    a = dev_set["aw"][f:f + 1].detach().clone().requires_grad_(True)
    go = dev_set["go"][f:f + 1]
    out = fn(v, shapes, lsi, l, a, 64)
    out.backward(go)
    torch.cuda.synchronize()
    ref_out, rv, rl, ra = O.core_pytorch_fwd_bwd(v.detach().cpu().double(), inp_cpu["shapes"], l.detach().cpu().double(),
                                                 a.detach().cpu().double(), go.cpu().double())
    e_gl, _, frac = O.grad_loc_errors(l.grad.float().cpu(), rl, l.detach().cpu(), wl.levels)
    errs = {"out": O.max_rel(out.detach().float().cpu(), ref_out), "grad_value": O.max_rel(v.grad.float().cpu(), rv),
            "grad_attn_weight": O.max_rel(a.grad.float().cpu(), ra), "grad_sampling_loc": e_gl}
    tol = {"out": 1e-5, "grad": 1e-4} if tdt == torch.float32 else {"out": 1e-2, "grad": 1e-2}
    ok = errs["out"] < tol["out"] and all(errs[k] < tol["grad"] for k in ("grad_value", "grad_attn_weight", "grad_sampling_loc"))
    return ok, errs, tol, frac


def run_ours(args, wl, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the operator has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    affinity = bind_to_gpu_numa_node(local_rank)   # before any pinned host buffer is allocated

    pkg = importlib.import_module("tce-rvos_b200")
    capi = importlib.import_module("tce-rvos_b200.capi")
    ext = pkg.load_extension()
    from tools import synthetic                    # seeded workload generator; the oracle is not imported before the timed regions

    tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
    N, S, Lq = wl.frames, wl.S, wl.S
    rep = (N + wl.base_frames - 1) // wl.base_frames

    def grow(t):                                   # base frames -> N frames of distinct device memory
        return t if rep == 1 else t.repeat(rep, *([1] * (t.dim() - 1)))[:N].contiguous()

    sets = []
    for i in range(wl.n_sets):
        inp = synthetic.make_inputs("model", wl.base_frames, wl.levels, M=M, D=D, P=P, seed=100 * rank + i, dtype=torch.float32)
        sets.append(dict(value=grow(inp["value"].to(dev, tdt)).requires_grad_(True), loc=grow(inp["loc"].to(dev)).requires_grad_(True),
                         aw=grow(inp["aw"].to(dev)).requires_grad_(True), go=grow(inp["grad_out"].to(dev, tdt))))
    shapes, lsi = inp["shapes"].to(dev), inp["lsi"].to(dev)
    host_set = inp                                                  # last set, for the e2e leg and the shard check
    issued, unmerged = count_line_updates(inp["loc"], wl.levels)
    line_updates, line_updates_unmerged = issued * (N / wl.base_frames), unmerged * (N / wl.base_frames)
    qh_per_step = N * Lq * M
    fn = pkg.MSDeformAttnFunction.apply

    def eager_step(i, ev=None):
        s = sets[i % wl.n_sets]
        out = fn(s["value"], shapes, lsi, s["loc"], s["aw"], 64)
        if ev is not None:
            ev[0].record()
        grads = torch.autograd.grad(out, (s["value"], s["loc"], s["aw"]), s["go"])
        if ev is not None:
            ev[1].record()
        return grads

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        eager_step(i)
    torch.cuda.synchronize()

    # ---- CUDA graphs of the whole step, one per input set (the step has no collective: the host only launches) ----
    graphs, launches_per_step = None, None
    c0 = ext.launch_count()
    eager_step(0)
    launches_per_step = ext.launch_count() - c0
    if not args.no_graph:
        graphs = []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(wl.n_sets):
                eager_step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for i in range(wl.n_sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep = eager_step(i)
            graphs.append((g, keep))
        for i in range(max(3, wl.n_sets)):
            graphs[i % wl.n_sets][0].replay()
        torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: exactly K steps, CUDA events on the launching (current) stream ----------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    if graphs is not None:
        for i in range(args.steps):
            graphs[i % wl.n_sets][0].replay()
    else:
        for i in range(args.steps):
            eager_step(i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * qh_per_step * args.steps / (ms_total * 1e-3) / 1e9

    # ---- the same K steps again, eagerly, with an event after each pass: per-kernel times for the roofline ----
    marks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):                     # torch.cuda.graph() empties the allocator cache: refill it before timing
        eager_step(i)
    torch.cuda.synchronize()
    x0.record()
    for i in range(args.steps):
        eager_step(i, marks[i])
    x1.record()
    torch.cuda.synchronize()
    eager_ms = x0.elapsed_time(x1) / args.steps
    fwd_ms, bwd_ms, prev = [], [], x0
    for a, b in marks:
        fwd_ms.append(prev.elapsed_time(a))
        bwd_ms.append(a.elapsed_time(b))
        prev = b

    # ---- post-roll: keep the same step running long enough for the clock sampler ----------------
    t_end = time.perf_counter() + 1.0
    i = 0
    while time.perf_counter() < t_end:
        if graphs is not None:
            graphs[i % wl.n_sets][0].replay()
        else:
            eager_step(i)
        i += 1
        if i % 50 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    sampler.stop()

    # ---- shard parity: every rank checks one of its frames against the oracle (outside the timed regions) ----
    ok, errs, tol, bfrac = shard_parity(wl, host_set, sets[0], shapes, lsi, fn, tdt)
    parity = torch.tensor([0.0 if ok else 1.0, errs["out"], errs["grad_value"], errs["grad_attn_weight"], errs["grad_sampling_loc"]],
                          device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(parity, op=dist.ReduceOp.MAX)
    parity = parity.tolist()

    # ---- e2e: HOST buffers through the C-ABI host call, copies inside the timed region ----------
    e2e_frames = N if wl.name == "config2" else min(N, 16)
    vdt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    e2e_rep = (e2e_frames + wl.base_frames - 1) // wl.base_frames

    def grow_host(t):
        return (t if e2e_rep == 1 else t.repeat(e2e_rep, *([1] * (t.dim() - 1))))[:e2e_frames].contiguous()

    pin = {k: grow_host(host_set[k].to(vdt if k in ("value", "grad_out") else torch.float32)).pin_memory()
           for k in ("value", "loc", "aw", "grad_out")}
    res = dict(out=torch.empty(e2e_frames, Lq, M * D, dtype=vdt).pin_memory(), gv=torch.empty_like(pin["value"]).pin_memory(),
               gl=torch.empty_like(pin["loc"]).pin_memory(), ga=torch.empty_like(pin["aw"]).pin_memory())
    sh = capi.shape_of(pin["value"], pin["loc"], L)
    ctx = capi.HostContext(local_rank)
    shp_h, lsi_h = host_set["shapes"], host_set["lsi"]

    def e2e_step():
        ctx.forward_backward(sh, pin["value"], shp_h, lsi_h, pin["loc"], pin["aw"], pin["grad_out"],
                             res["out"], res["gv"], res["gl"], res["ga"])

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        e2e_step()
    launches_e0 = ext.launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()                                                  # blocking: returns with results on the host
    e2e_s = time.perf_counter() - t0
    barrier()
    launches_e2e = ext.launch_count() - launches_e0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * e2e_frames * Lq * M * e2e_steps / e2e_s / 1e9
    h2d = sum(pin[k].numel() * pin[k].element_size() for k in pin)
    d2h = sum(res[k].numel() * res[k].element_size() for k in res)
    ctx.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant pass -----------------------------------------------------------
    lib = capi.load()
    shape = capi.make_shape(N, S, Lq, M, D, L, P, capi.dtype_code(tdt), capi.MSDA_F32)
    bytes_fwd, bytes_bwd = lib.msda_algorithmic_bytes(shape, 0), lib.msda_algorithmic_bytes(shape, 1)
    peak, peak_src = measured_peak_gbs()
    f_us, b_us = statistics.mean(fwd_ms) * 1e3, statistics.mean(bwd_ms) * 1e3
    step_us = ms_total * 1e3 / args.steps
    dominant = "backward" if b_us >= f_us else "forward"
    dom_bytes, dom_us = (bytes_bwd, b_us) if dominant == "backward" else (bytes_fwd, f_us)
    achieved = dom_bytes / (dom_us * 1e-6) / 1e9
    roofline = {"bound": "hbm", "kernel": lib.msda_kernel_name(shape, 1 if dominant == "backward" else 0).decode(),
                "pass": dominant + (" (grad_value zero-fill + scatter kernel)" if dominant == "backward" else ""),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_us": dom_us,
                "timed_by": "CUDA events after each pass in the eager re-run of the same steps that follows the graph-replayed "
                            "timed region" if graphs is not None else "CUDA events after each pass inside the timed region",
                "passes": {"forward": {"us": f_us, "algorithmic_bytes": bytes_fwd, "GB/s": bytes_fwd / f_us / 1e3,
                                       "frac": bytes_fwd / f_us / 1e3 / peak},
                           "backward": {"us": b_us, "algorithmic_bytes": bytes_bwd, "GB/s": bytes_bwd / b_us / 1e3,
                                        "frac": bytes_bwd / b_us / 1e3 / peak},
                           "fwd+bwd": {"us": step_us, "GB/s": (bytes_fwd + bytes_bwd) / step_us / 1e3,
                                       "frac": (bytes_fwd + bytes_bwd) / step_us / 1e3 / peak}}}
    # the unit that binds the backward (profiles/r02_bwd_split.md): the L2's reduction throughput — with the corner loads
    # removed the (unmerged) pass still takes 93 % of its time, with the reductions removed 46 %; the kernel therefore
    # merges the duplicate corners of consecutive points before reducing them (profiles/r02_merge.md)
    roofline["l2_reduction"] = {
        "what": "backward pass: 128-byte grad_value line updates (red.global.add.v4.f32) per second against the rate the chip's L2 "
                "sustains in a microbenchmark on this pool from 74 SMs upwards (profiles/r01_microbench_bulk_red.txt); "
                "split experiment: profiles/r02_bwd_split.md",
        "line_updates_per_launch": line_updates, "line_updates_before_merging": line_updates_unmerged,
        "achieved_Glines_s": line_updates / (b_us * 1e-6) / 1e9,
        "peak_Glines_s": L2_REDUCTION_PEAK_GLINES, "frac": line_updates / (b_us * 1e-6) / 1e9 / L2_REDUCTION_PEAK_GLINES,
        "payload_TB_s": line_updates * 128 / (b_us * 1e-6) / 1e12}
    traffic_path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_path) and wl.name == "config2":
        try:
            with open(traffic_path) as f:
                roofline["traffic"] = json.load(f).get(roofline["kernel"])
        except Exception:
            pass

    # ---- CPU baseline (rank 0, N=1 only): fixed sample ------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cv, csec, threads, kind, sample = time_cpu(wl, 1, 5 if wl.name == "config2" else 2)
        cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "s_per_step": csec}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": wl.config(args.dtype, world),
            "timed_region": ("CUDA-graph replays of the whole step (MSDeformAttnFunction.apply forward + autograd backward), one graph "
                             "per input set" if graphs is not None else "eager steps"),
            "ms_per_step_eager": eager_ms,
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3, "frames_per_gpu": e2e_frames,
                    "api": "msda_forward_backward_host (C ABI, pinned host buffers in and out, per-frame "
                           "copy/compute pipeline on 3 streams); host wall clock around the blocking calls"},
            "shard_parity": {"ok": parity[0] == 0.0, "ranks_checked": world, "frame": "frame 0 of every rank's input set 0",
                             "max_err_over_ranks": {"out": parity[1], "grad_value": parity[2], "grad_attn_weight": parity[3],
                                                    "grad_sampling_loc": parity[4]},
                             "tolerance": tol, "metric": "max|a-b| / max|b| against the fp64 oracle; grad_sampling_loc excludes points "
                                                         "within 1e-4 px of a bilinear cell boundary", "boundary_fraction": bfrac},
            "host_affinity_cpus": affinity,
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
            "gpu_launches_e2e": int(launches_e2e),
            "clocks": sampler.summary()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# whole-model workloads (BASELINE.json configs[2], configs[3]) on the staged, unmodified reference model
# ------------------------------------------------------------------------------------------------
def run_model(args, rank, local_rank, world):
    cfg = 3 if args.workload == "model3" else 4
    name = f"BASELINE.json configs[{cfg - 1}]: " + ("TCE-RVOS ResNet-50 random-init inference, 36-frame 360x640 clip, 6+6 layers"
                                                    if cfg == 3 else "TCE-RVOS Video-Swin-T random-init training step, 2 clips x 5 frames per GPU, DDP")
    if not os.path.isdir(STAGED):
        if rank == 0:
            print(json.dumps({"metric": "rvos_frames_per_s", "workload": args.workload, "config": {"workload": name},
                              "unavailable": "baseline/_ref/tce_rvos_src is absent: stage the reference's python tree with "
                                             "`python tools/model_dropin.py --stage` in the build container"}), flush=True)
        return
    import torch
    from tools import model_dropin
    sampler = ClockSampler(local_rank)
    sampler.start()
    argv = ["--config", str(cfg), "--op", "reference" if args.impl == "reference" else "b200",
            "--steps", str(args.steps), "--warmup", str(max(args.warmup, 2))]
    r = model_dropin.run(argv)
    sampler.stop()
    if rank != 0 or r is None:
        return
    line = {"metric": "rvos_frames_per_s", "value": r["frames_per_s"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 2), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference-cuda-kernels" if args.impl == "reference" else "ours",
            "config": {"workload": name, "backbone": r["backbone"], "enc_dec_layers": r["enc_dec_layers"],
                       "frames_per_clip": r["frames_per_clip"], "clips_per_gpu": r["clips_per_gpu"], "input": r["input"],
                       "mode": r["mode"], "weights": r["weights"],
                       "parallelism": "DDP, NCCL gradient all-reduce (main.py:67-72)" if cfg == 4 else "one clip per GPU, replicas (inference_ytvos.py:101-113)"},
            "operator": {"extension": os.path.relpath(r["op_extension"], ROOT) if os.path.isabs(str(r["op_extension"])) else r["op_extension"],
                         "calls_per_step": r["op_calls_per_step"], "ms_per_step": r["op_ms_per_step"],
                         "kernel_launches_per_step": r["b200_kernel_launches_per_step"]},
            "gpu_launches": int((r["b200_kernel_launches_per_step"] or 0) * args.steps),
            "result_checksum": r["result_checksum"], "clocks": sampler.summary()}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--workload", default="config2", choices=["config2", "sweep", "model3", "model4"])
    ap.add_argument("--frames", type=int, default=None, help="sweep: frames per GPU (64..1024)")
    ap.add_argument("--no-graph", action="store_true", help="time eager steps instead of CUDA-graph replays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.workload in ("model3", "model4"):
        args.steps = args.steps or 5
        args.warmup = 2 if args.warmup is None else args.warmup
        return run_model(args, rank, local_rank, world)
    wl = Workload(args.workload, args.frames)
    if args.steps is None:
        args.steps = 200 if wl.name == "config2" else max(5, min(50, 3200 // wl.frames))
    if args.warmup is None:
        args.warmup = 10 if wl.name == "config2" else 3
    args.warmup = max(args.warmup, 3 if args.impl == "ours" else 1)
    if args.impl == "reference":
        run_reference(args, wl, rank)
    else:
        run_ours(args, wl, rank, local_rank, world)


if __name__ == "__main__":
    main()
