"""bench.py pieces that run without a GPU: the reference arm (C oracle on the host cores, same records as the GPU
arm, CPU per-time-step latency leg) and the sharding of the record blocks over the ranks."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scenarios", "8",
                          "--steps", "2", "--warmup", "1", "--no-cache", "--sim-steps", "6", "--gen-workers", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "vehicle-plans/sec" and line["unit"] == "plans/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1
    n = line["config"]["searches_per_step_per_gpu"]
    assert n == 8 * 6 * 20 and "same records in both arms" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and f"all {n} search records" in cb["sample"]
    lat = cb["latency_ms_per_timestep"]
    assert lat["n"] == 2 * 8 * 6 and 0 < lat["p50"] <= lat["p99"] <= lat["max"]
    assert line["e2e"] == {"value": line["value"], "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(line["value"] - n * 2 / (line["ms_per_step"] * 2e-3)) < 1e-6 * line["value"]


def test_rank_blocks_weak_and_strong():
    sys.path.insert(0, ROOT)
    import bench

    class A:
        scenarios, scenarios_total = 512, 0
    assert bench.rank_blocks(A, 0, 1) == list(range(64)) and bench.rank_blocks(A, 3, 8) == list(range(192, 256))
    A.scenarios_total = 4096
    parts = [bench.rank_blocks(A, r, 8) for r in range(8)]
    assert sorted(b for p in parts for b in p) == list(range(512)) and all(len(p) == 64 for p in parts)
    assert parts[1][:3] == [1, 9, 17]                                   # block-cyclic
    cfg = bench.workload_config(type("X", (), {"scenarios_total": 4096, "scenarios": 512, "sim_steps": 35, "optimizer": "graph",
                                               "vehicles": 20, "mpa": "triple_speed", "mcts_expansions": 250}), 8, 358400, 6)
    assert "4096 scenarios in total" in cfg["workload"] and cfg["searches_per_step_per_gpu"] == 358400
