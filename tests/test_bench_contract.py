"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm (the reference's CPU path timed alone),
the orbit clock, and the strip helpers bench.py builds on."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference`: one JSON line, impl/metric/unit/config of the B200 arm, cpu_baseline describing the run,
    e2e with zero copy bytes; bounded sample."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "config1", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ReSTIR DI Mpixel/s" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "config1" and d["config"]["resolution"] == [800, 800]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["value"] - 0.64 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_orbit_clock_sweeps_the_60_frame_arc_back_and_forth():
    sys.path.insert(0, ROOT)
    import bench
    idx = [bench.orbit_index(k) for k in range(240)]
    assert idx[:60] == list(range(60)) and idx[60:120] == list(range(59, -1, -1)) and idx[120:] == idx[:120]
    assert max(abs(a - b) for a, b in zip(idx, idx[1:])) <= 1          # consecutive frames stay temporally coherent
    w = bench.window_frames(5, 20)                                       # the driver's --warmup 5 --steps 20: cuts for the timed stretch
    assert 1 <= len(w) <= 6 and w[0] == 5 and all(5 <= k < 25 for k in w) and w == sorted(w)
    assert bench.window_frames(0, 1) == [0] and len(bench.window_frames(51, 26)) <= 6


def test_gi_side_measurement_never_fails_the_bench(monkeypatch):
    """bench.py's `restir_gi` block comes from scripts/gi_bench.py in a child process: its lines are folded into one dict; without a GPU
    (or with a child that dies or hangs) it degrades to {"error": ...} instead of raising."""
    sys.path.insert(0, ROOT)
    import bench

    first = bench.gi_side_measurement(120)        # without a GPU the child exits with a RestirError; with one it measures
    assert set(first) == {"error"} or first["ms_per_frame"]["ray_queues"] > 0
    rows = [l for l in open(os.path.join(ROOT, "profiles", "r02_c35_gi_bench.jsonl")) if "config4_1080p" in l]

    class Done:
        stdout, stderr = "import noise\n" + "".join(rows), ""

    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: Done)
    d = bench.gi_side_measurement()
    assert d["workload"] == "config4_1080p" and d["trace_depth"] == 3 and set(d["ms_per_frame"]) == {"ray_queues", "staged"}
    assert 0 < d["ms_per_frame"]["ray_queues"] < d["ms_per_frame"]["staged"]

    def boom(*a, **k):
        raise subprocess.TimeoutExpired("gi_bench", 1)

    monkeypatch.setattr(bench.subprocess, "run", boom)
    assert "error" in bench.gi_side_measurement()
