"""bench.py contract on the CPU box: the reference arm (the reference's CPU implementation of the path, no GPU, no
extension of this repo), its JSON keys, and the rank rule under torchrun (only rank 0 works and prints)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None); env.pop("LOCAL_RANK", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *args],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "msdeformattn_fwd_bwd_query_heads_per_s"
    assert line["unit"] == "Gquery-heads/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["steps"] == 1 and line["value"] > 0 and line["gpu_launches"] == 0
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["value"] == line["value"]
    assert "1 frame x all 4820 queries" in cpu["sample"]          # the FIXED sample (reproducible ratio)
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in line["config"]["workload"] and line["config"]["frames_per_gpu"] == 10
    staged = (ROOT / "baseline" / "_ref" / "tce_rvos_src" / "models" / "ops" / "functions" / "ms_deform_attn_func.py").exists()
    assert cpu["kind"] == ("reference" if staged else "port")


def test_reference_arm_other_ranks_do_nothing():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_sweep_workload_names_itself():
    r = _run(args=("--workload", "sweep", "--frames", "128"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert "configs[4]" in line["config"]["workload"] and line["config"]["frames_per_gpu"] == 128 and line["config"]["S"] == 9520
