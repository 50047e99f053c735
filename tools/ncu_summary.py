"""Turn ncu reports (gpurun_out/*.ncu-rep, launches CSV) into the tracked summaries under profiles/.

    python tools/ncu_summary.py <round-tag> <launches.csv> <name>=<report.ncu-rep> ...
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
]


def raw_rows(rep):
    if rep.endswith(".csv"):                      # a raw page exported on the GPU box (tools/final_validation.sh)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    tag, launches = sys.argv[1], sys.argv[2]
    os.makedirs(PROF, exist_ok=True)
    traffic_path = os.path.join(PROF, "dram_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}

    # ---- launch list ----------------------------------------------------------------------------
    if os.path.exists(launches):
        rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
        hdr = rows[0]
        first_n = int(os.environ.get("FIRST_N", "110"))   # (warm-up + steps) x 2 kernels of the device-timed loop
        rows = rows[: 1 + first_n]
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        per = OrderedDict()
        for r in rows[1:]:
            try:
                per.setdefault(r[ik], []).append(float(r[iv].replace(",", "")))
            except ValueError:
                pass
        total = sum(sum(v) for v in per.values())
        with open(os.path.join(PROF, f"{tag}_launches_summary.md"), "w") as f:
            cmd = os.environ.get("BENCH_CMD", "python bench.py --steps 20 --warmup 5 --no-cpu-baseline")
            f.write(f"# {tag}: ncu launch list of `{cmd}`\n\n")
            f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (per-launch times are cold-cache and "
                    f"serialised: compare SHARES).  The first {first_n} launches are summarised: the warm-up steps, the CUDA-graph "
                    "replays of the timed region (ncu profiles the kernel nodes of a replayed graph one by one) and the instrumented "
                    "eager re-run, 2 kernels per step (the grad_value zero-fill is a cudaMemsetAsync, not a kernel); the post-roll "
                    "and the e2e leg that follow launch the same two kernels.  Full list: `" + f"{tag}_launches.csv`.\n\n")
            f.write("| kernel | launches | mean us | share of captured GPU time |\n|---|---|---|---|\n")
            for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
                f.write(f"| `{k[:110]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / total:.1f}% |\n")
        with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["id", "kernel", "gpu__time_duration.sum [ns]"])
            for r in rows[1:]:
                w.writerow([r[0], r[ik][:140], r[iv]])

    # ---- full captures ----------------------------------------------------------------------------
    for spec in sys.argv[3:]:
        name, rep = spec.split("=")
        hdr, units, data = raw_rows(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        with open(os.path.join(PROF, f"{tag}_ncu_{name}.md"), "w") as f:
            f.write(f"# {tag}: ncu --set full, `{name}`\n\nSource report: `{os.path.basename(rep)}` (not tracked; "
                    f"captured with `--clock-control none --import-source on` on B200 through gpurun).\n\n")
            for r in data:
                kname = r[idx["Kernel Name"]]
                f.write(f"## `{kname[:150]}`\n\n| metric | value | unit |\n|---|---|---|\n")
                for k in KEYS:
                    if k in idx and r[idx[k]] != "":
                        f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
                # derived: sectors per request (north_star asks for it), achieved DRAM GB/s
                def num(k):
                    try:
                        return float(r[idx[k]].replace(",", ""))
                    except Exception:
                        return None
                for op in ("ld", "red"):
                    rq, sc = num(f"l1tex__t_requests_pipe_lsu_mem_global_op_{op}.sum"), num(f"l1tex__t_sectors_pipe_lsu_mem_global_op_{op}.sum")
                    if rq and sc:
                        f.write(f"| derived: sectors per request, global {op} | {sc / rq:.2f} | sector/request |\n")
                try:
                    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
                    b = num("dram__bytes_read.sum") * scale[units[idx["dram__bytes_read.sum"]]] + \
                        num("dram__bytes_write.sum") * scale[units[idx["dram__bytes_write.sum"]]]
                    t_us = num("gpu__time_duration.sum") * {"us": 1, "ms": 1e3, "ns": 1e-3}[units[idx["gpu__time_duration.sum"]]]
                    f.write(f"| derived: achieved DRAM throughput | {b / t_us / 1e3:.0f} | GB/s |\n")
                except Exception:
                    pass
                f.write("\n")
                try:
                    b = float(r[idx["dram__bytes_read.sum"]]) + float(r[idx["dram__bytes_write.sum"]])
                    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
                    ur, uw = units[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_write.sum"]]
                    b = float(r[idx["dram__bytes_read.sum"]]) * scale[ur] + float(r[idx["dram__bytes_write.sum"]]) * scale[uw]
                    traffic[name] = int(b)
                except Exception:
                    pass
    json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
