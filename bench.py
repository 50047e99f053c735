#!/usr/bin/env python
"""bench.py — throughput of the MSDeformAttn hot path on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): one step = MSDeformAttn forward + backward over one batch of
synthetic encoder-shaped input at TCE-RVOS's training shape — B*T = 2x5 = 10 frames per GPU,
levels 45x80 / 23x40 / 12x20 / 6x10 (S = 4820), Lq = S, 8 heads x 32 channels, 4 points, fp32 —
through the reference-facing MSDeformAttnFunction.  Metric: query-heads (N*Lq*M) per second.

* `value`      inputs resident in HBM, CUDA-event timing on the launching stream, max over ranks.
* `e2e`        same step through the C-ABI host-buffer call (msda_forward_backward_host): pinned
               HOST inputs, host<->device copies inside the timed region.
* `roofline`   the dominant pass (backward: zero-fill + scatter kernel), algorithmic bytes / its
               CUDA-event time, against MEASURED_PEAKS.json's copy bandwidth.
* `cpu_baseline` the oracle's restatement of the reference's CPU path
               (ms_deform_attn_core_pytorch + autograd), timed on this host on a bounded sample.
* `--impl reference`  times that CPU path alone (all host threads), same metric and config.

Frames shard across ranks with no collective (weak scaling: every rank owns 10 frames).
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

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LEVELS_A = [(45, 80), (23, 40), (12, 20), (6, 10)]
FRAMES_PER_GPU = 10
M, D, L, P = 8, 32, 4, 4
METRIC = "msdeformattn_fwd_bwd_query_heads_per_s"
UNIT = "Gquery-heads/s"
N_INPUT_SETS = 3                      # rotated so that no step finds its inputs in L2
CPU_SAMPLE_FRAMES = 2


def workload_config(dtype, n_gpus):
    S = sum(h * w for h, w in LEVELS_A)
    return {"workload": "BASELINE.json configs[1]: MSDeformAttn fwd+bwd, B*T=2x5 frames/GPU, levels "
                        "45x80/23x40/12x20/6x10 (S=4820), Lq=S, 8 heads x 32 ch, 4 levels x 4 points, "
                        "model-like sampling locations (SURVEY.md 8d-2)",
            "frames_per_gpu": FRAMES_PER_GPU, "S": S, "Lq": S, "heads": M, "channels": D, "levels": L,
            "points": P, "value_dtype": dtype, "loc_dtype": "f32", "sharding": f"frames x {n_gpus} ranks, no collective",
            "l2_policy": f"{N_INPUT_SETS} rotating input sets, each 173 MB fwd / 469 MB fwd+bwd > 126 MB L2"}


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
        self.samples, self.reasons, self.max_mhz = [], set(), None
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
# reference arm / CPU baseline: the oracle's port of the reference's CPU implementation
# ------------------------------------------------------------------------------------------------
def cpu_step_inputs(frames, lq=None, seed=0):
    import torch
    from tools import synthetic
    inp = synthetic.make_inputs("model", frames, LEVELS_A, M=M, D=D, P=P, seed=seed, dtype=torch.float32)
    if lq is not None and lq < inp["loc"].shape[1]:
        # a bounded sample: the first lq queries of every frame (value keeps its full size)
        for k in ("loc", "aw", "grad_out"):
            inp[k] = inp[k][:, :lq].contiguous()
    return inp


def cpu_step(inp):
    from oracle import msda_oracle as O
    return O.core_pytorch_fwd_bwd(inp["value"], inp["shapes"], inp["loc"], inp["aw"], inp["grad_out"])


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every core it is allowed."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    if torch.get_num_threads() != n:
        torch.set_num_threads(n)
    return n


def time_cpu(frames, warmup, steps, lq=None):
    import torch
    use_all_host_threads()
    inp = cpu_step_inputs(frames, lq)
    for _ in range(warmup):
        cpu_step(inp)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(inp)
    dt = time.perf_counter() - t0
    qh = frames * inp["loc"].shape[1] * M
    return qh * steps / dt / 1e9, dt / steps, torch.get_num_threads()


def bounded_cpu_sample(total_passes, budget_s=100.0):
    """Pick (frames, Lq) of one CPU step so that `total_passes` steps fit the time budget."""
    S = sum(h * w for h, w in LEVELS_A)
    _, probe_s, _ = time_cpu(1, 0, 1, lq=S // 4)
    rate = (S // 4) * M / probe_s                               # query-heads per second, roughly size-independent
    per_step_qh = rate * budget_s / max(total_passes, 1)
    if per_step_qh >= CPU_SAMPLE_FRAMES * S * M:
        return CPU_SAMPLE_FRAMES, S
    if per_step_qh >= S * M:
        return 1, S
    return 1, int(max(256, min(S, per_step_qh / M)))


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (pure-PyTorch
    grid_sample formulation + autograd), via the oracle port, all host threads, rank 0 only."""
    if rank != 0:
        return
    import torch
    frames, lq = bounded_cpu_sample(args.steps + args.warmup)
    value, sec, threads = time_cpu(frames, args.warmup, args.steps, lq=lq)
    cfg = workload_config("f32", 1)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"per step {frames} of the {FRAMES_PER_GPU} frames x the first {lq} of {sum(h * w for h, w in LEVELS_A)} "
                                       f"queries of the workload (same S, heads, points), fwd + autograd bwd of the oracle's "
                                       f"port of ms_deform_attn_core_pytorch, torch {torch.__version__} CPU, "
                                       f"os.cpu_count()={os.cpu_count()}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def count_line_updates(loc, levels):
    """Number of 128-byte grad_value line updates one backward issues: one per (query, head, point, corner)
    whose corner lies inside its level (the reference's gating, ms_deform_im2col_cuda.cuh:441 and :57-79)."""
    import torch
    total = 0
    for l, (H, W) in enumerate(levels):
        x = loc[:, :, :, l, :, 0] * W - 0.5
        y = loc[:, :, :, l, :, 1] * H - 0.5
        inside = (x > -1) & (y > -1) & (x < W) & (y < H)
        x0, y0 = torch.floor(x), torch.floor(y)
        for dy in (0, 1):
            for dx in (0, 1):
                ok = inside & (x0 + dx >= 0) & (x0 + dx <= W - 1) & (y0 + dy >= 0) & (y0 + dy <= H - 1)
                total += int(ok.sum())
    return total


# measured on this pool (profiles/r01_microbench_bulk_red.txt, r01_microbench_l1_shapes.txt): the chip sustains
# 47.5-49.4 G updates/s of 128-byte lines through red.global.add.v4.f32 (or TMA bulk reductions) = ~6.1 TB/s
L2_REDUCTION_PEAK_GLINES = 47.5


def bind_to_gpu_numa_node(index):
    """The e2e leg moves 173 MB each way per step through pinned host memory.  torchrun does not bind
    ranks to the socket their GPU hangs off, so with several ranks the copies cross the inter-socket
    link.  Pin this process (and therefore the first touch of its pinned buffers) to the CPUs NVML
    reports as local to the GPU.  Returns the number of CPUs in the set, or None when unavailable."""
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


def run_ours(args, rank, local_rank, world):
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
    from tools import synthetic                    # seeded workload generator; the oracle is not imported on this arm

    tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
    sets = []
    for i in range(N_INPUT_SETS):
        inp = synthetic.make_inputs("model", FRAMES_PER_GPU, LEVELS_A, M=M, D=D, P=P, seed=100 * rank + i, dtype=torch.float32)
        sets.append(dict(value=inp["value"].to(dev, tdt).requires_grad_(True), loc=inp["loc"].to(dev).requires_grad_(True),
                         aw=inp["aw"].to(dev).requires_grad_(True), go=inp["grad_out"].to(dev, tdt)))
    shapes, lsi = inp["shapes"].to(dev), inp["lsi"].to(dev)
    host_set = inp                                                  # last set, for the e2e leg
    line_updates = count_line_updates(sets[0]["loc"].detach(), LEVELS_A)   # what bounds the backward
    N, S, Lq = FRAMES_PER_GPU, inp["value"].shape[1], inp["loc"].shape[1]
    qh_per_step = N * Lq * M
    fn = pkg.MSDeformAttnFunction.apply

    def step(i, ev=None):
        s = sets[i % N_INPUT_SETS]
        s["value"].grad = s["loc"].grad = s["aw"].grad = None
        out = fn(s["value"], shapes, lsi, s["loc"], s["aw"], 64)
        if ev is not None:
            ev[0].record()
        out.backward(s["go"])
        if ev is not None:
            ev[1].record()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step(i)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: exactly K steps, CUDA events on the launching (current) stream ----------
    marks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ext.launch_count()
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i, marks[i])
    e1.record()
    barrier()
    launches = ext.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    # per-pass split from the same timed region: start -> after forward -> after backward
    fwd_ms, bwd_ms = [], []
    prev = e0
    for a, b in marks:
        fwd_ms.append(prev.elapsed_time(a))
        bwd_ms.append(a.elapsed_time(b))
        prev = b
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * qh_per_step * args.steps / (ms_total * 1e-3) / 1e9

    # ---- post-roll: keep the same step running long enough for the clock sampler ----------------
    t_end = time.perf_counter() + 1.0
    i = 0
    while time.perf_counter() < t_end:
        step(i); i += 1
        if i % 50 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    sampler.stop()

    # ---- e2e: HOST buffers through the C-ABI host call, copies inside the timed region ----------
    vdt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    pin = {k: host_set[k].to(vdt if k in ("value", "grad_out") else torch.float32).pin_memory()
           for k in ("value", "loc", "aw", "grad_out")}
    res = dict(out=torch.empty(N, Lq, M * D, dtype=vdt).pin_memory(), gv=torch.empty_like(pin["value"]).pin_memory(),
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
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * qh_per_step * e2e_steps / e2e_s / 1e9
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
    dominant = "backward" if b_us >= f_us else "forward"
    dom_bytes, dom_us = (bytes_bwd, b_us) if dominant == "backward" else (bytes_fwd, f_us)
    achieved = dom_bytes / (dom_us * 1e-6) / 1e9
    roofline = {"bound": "hbm", "kernel": lib.msda_kernel_name(shape, 1 if dominant == "backward" else 0).decode(),
                "pass": dominant + (" (grad_value zero-fill + scatter kernel)" if dominant == "backward" else ""),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_us": dom_us,
                "passes": {"forward": {"us": f_us, "algorithmic_bytes": bytes_fwd, "GB/s": bytes_fwd / f_us / 1e3,
                                       "frac": bytes_fwd / f_us / 1e3 / peak},
                           "backward": {"us": b_us, "algorithmic_bytes": bytes_bwd, "GB/s": bytes_bwd / b_us / 1e3,
                                        "frac": bytes_bwd / b_us / 1e3 / peak},
                           "fwd+bwd": {"us": ms_total * 1e3 / args.steps,
                                       "GB/s": (bytes_fwd + bytes_bwd) / (ms_total * 1e3 / args.steps) / 1e3,
                                       "frac": (bytes_fwd + bytes_bwd) / (ms_total * 1e3 / args.steps) / 1e3 / peak}}}
    # the bound that actually limits the backward: L2 reduction throughput (informative, next to the HBM figure)
    roofline["l2_reduction"] = {
        "what": "backward pass: 128-byte grad_value line updates (red.global.add.v4.f32) per second against the rate the "
                "L2 sustains in a microbenchmark on this pool (profiles/r01_microbench_bulk_red.txt)",
        "line_updates_per_launch": line_updates, "achieved_Glines_s": line_updates / (b_us * 1e-6) / 1e9,
        "peak_Glines_s": L2_REDUCTION_PEAK_GLINES, "frac": line_updates / (b_us * 1e-6) / 1e9 / L2_REDUCTION_PEAK_GLINES,
        "payload_TB_s": line_updates * 128 / (b_us * 1e-6) / 1e12}
    traffic_path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_path):
        try:
            with open(traffic_path) as f:
                roofline["traffic"] = json.load(f).get(roofline["kernel"])
        except Exception:
            pass

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cv, csec, threads = time_cpu(CPU_SAMPLE_FRAMES, 1, 2)
        cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{CPU_SAMPLE_FRAMES} of the {FRAMES_PER_GPU} frames (same S, Lq, heads, points), fwd + autograd "
                         f"bwd of the oracle's port of ms_deform_attn_core_pytorch, 1 warm-up + 2 timed passes, "
                         f"{csec:.2f} s/pass, os.cpu_count()={os.cpu_count()}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args.dtype, world),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "api": "msda_forward_backward_host (C ABI, pinned host buffers in and out, per-frame "
                           "copy/compute pipeline on 3 streams); host wall clock around the blocking calls"},
            "host_affinity_cpus": affinity,
            "gpu_launches": int(launches), "gpu_launches_e2e": int(ext.launch_count() - launches_e0),
            "clocks": sampler.summary()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "ours" else 1)

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
