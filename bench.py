#!/usr/bin/env python3
"""bench.py — vehicle-plans/sec of the MPA graph-search optimizer on B200.

Workload at N=1 (BASELINE.json configs[1]): CPM Lab road network, 20 vehicles,
coloring-based prioritisation, triple_speed MPA, Hp 6, InterX checker, 35 time
steps per scenario, R scenarios planned concurrently ("parallel_threads
equivalent on 1 B200").  Scenarios are rolled out closed loop once (untimed) with
the GPU planner so that every (scenario, step, vehicle) search record is fixed;
one bench "step" = one pass of the hot path over all records of the rank.

  value     : records already resident in HBM, search kernel timed with CUDA
              events on the library's stream, L2 flushed between steps.
  e2e       : same records through the C-ABI call pdmpc_plan_batch with HOST
              (pinned) buffers: H2D staging + kernels + D2H results inside the
              timed region.
  roofline  : algorithmic bytes of the search kernel (SURVEY.md §8d formula from
              the kernel's own counters) / its CUDA-event time vs measured HBM peak;
              plus the FP64-pipe figure, which is the bound that actually bites.
  cpu_baseline : the C oracle (a port of the reference; MATLAB is not available)
              timed on this box's host cores on a bounded sample of the same records.

`--impl reference` times that CPU implementation alone (all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenarios", type=int, default=int(os.environ.get("PDMPC_BENCH_SCENARIOS", "256")),
                    help="scenarios per GPU (weak scaling)")
    ap.add_argument("--gen-workers", type=int, default=0,
                    help="processes that roll the scenarios out (0 = auto: host cores / ranks, at most 12)")
    ap.add_argument("--sim-steps", type=int, default=35)
    ap.add_argument("--vehicles", type=int, default=20)
    ap.add_argument("--mpa", default="triple_speed")
    ap.add_argument("--no-cache", action="store_true", help="do not read/write the record cache under build/")
    ap.add_argument("--cpu-sample", type=int, default=0, help="searches in the CPU sample (0 = auto)")
    ap.add_argument("--pipeline-chunks", type=int, default=0,
                    help="chunks of the e2e call's copy/search pipeline (0 = library default, 1 = off)")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), \
        int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _roll_chunk(job):
    """Worker process: roll a chunk of scenarios out closed loop with its OWN GPU planner."""
    dev, mpa_type, vehicles, sim_steps, seeds, out_path = job
    from pdmpc_b200 import capi, scenario
    from pdmpc_b200.mpa import get_mpa
    from pdmpc_b200.records import SearchBatch
    mpa = get_mpa(mpa_type, non_convex=True)
    planner = capi.Planner(dev)
    planner.upload_mpa(mpa)
    from pdmpc_b200.records import TimestepDeps
    batches, levels, timesteps = [], [], []
    for si, s in enumerate(seeds):
        sc = scenario.commonroad_scenario(mpa, vehicles, seed=s)
        runner = scenario.ScenarioRunner(sc, planner.plan_batch)
        if si < 8:   # level structure + one-call inputs of the first scenarios (per-time-step latency replay)
            for _ in range(sim_steps):
                iters, preds, fbs = runner.timestep_inputs()
                timesteps.append((si * 100000 + runner.k + 1,
                                  SearchBatch.from_iters(iters, mpa.Hp, sc.checker, mpa.dt_seconds),
                                  TimestepDeps.build(preds, [f[0] for f in fbs], mpa.Hp)))
                runner.step()
            recs = runner.records
            levels.extend((si * 100000 + r.step, r.level, r.batch.n) for r in recs)
        else:
            recs = runner.run(sim_steps)
        batches.extend(r.batch for r in recs)
    planner.close()
    SearchBatch.concat(batches).save(out_path)
    return out_path, levels, timesteps


def build_records(dev, mpa_type, n_scen: int, seed0: int, vehicles: int, sim_steps: int, cache: str = "",
                  workers: int = 1):
    """Closed-loop roll-out of n_scen road-network scenarios with the GPU planner (untimed),
    spread over `workers` processes (the scenario logic around the planner is host-side
    Python).  Returns (flat batch of every search record, [(step, level, n_searches)] of the
    first scenario, whose records come first in the batch).  Cached under build/
    (git-ignored) so that profiler runs of the same command skip the generation launches."""
    from pdmpc_b200.records import SearchBatch
    import pickle
    if cache and os.path.exists(cache) and os.path.exists(cache + ".levels.npy") and os.path.exists(cache + ".ts.pkl"):
        return SearchBatch.load(cache), np.load(cache + ".levels.npy"), pickle.load(open(cache + ".ts.pkl", "rb"))
    import multiprocessing as mp
    import tempfile
    workers = max(1, min(workers, n_scen))
    tmp = tempfile.mkdtemp(prefix="pdmpc_gen_")
    seeds = [seed0 + s for s in range(n_scen)]
    per = (n_scen + workers - 1) // workers
    jobs = [(dev, mpa_type, vehicles, sim_steps, seeds[i * per:(i + 1) * per], os.path.join(tmp, f"c{i}.npz"))
            for i in range(workers) if seeds[i * per:(i + 1) * per]]
    if len(jobs) == 1:
        results = [_roll_chunk(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(len(jobs)) as pool:
            results = pool.map(_roll_chunk, jobs)
    batch = SearchBatch.concat([SearchBatch.load(pth) for pth, _, _ in results])
    levels = np.array(results[0][1], dtype=np.int64)
    timesteps = results[0][2]
    for pth, _, _ in results:
        os.remove(pth)
    os.rmdir(tmp)
    if cache:
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        batch.save(cache)
        np.save(cache + ".levels.npy", levels)
        pickle.dump(timesteps, open(cache + ".ts.pkl", "wb"))
    return batch, levels, timesteps


def algorithmic_bytes(batch, stats, Hp: int) -> float:
    """B_in + 80*M + 80*P + B_out per search, summed over the batch.
    The per-pop obstacle re-read term is counted separately (obstacle columns are
    L1/L2 resident, not HBM traffic)."""
    b_in = batch.input_bytes()
    b_out = batch.n * (8 * (Hp * (4 + 2 * 7)) + 16)
    return float(b_in + 80 * stats.total_nodes + 80 * stats.total_pops + b_out)


def fp64_ops(batch, stats, Hp: int) -> float:
    """FP64 instructions (mul/add/sqrt, no FMA) the algorithm needs: per pop the
    sin/cos (2 x ~45) and shape placement (~6 pts x 2 shapes x 6), per obstacle
    column tested ~6 edges x 5 ops (C1 pass of InterX), per created node pose+g+h
    (~12 + 8 + 9*(Hp-1)/2)."""
    return float(stats.total_pops * (90 + 72) + stats.total_obstacle_cols * 30 +
                 stats.total_nodes * (20 + 9 * (Hp - 1) / 2))


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path (oracle port; the
    MATLAB reference cannot run here) on all host threads, bounded sample."""
    if rank != 0:
        return
    from oracle import oracle_py
    from pdmpc_b200 import scenario
    from pdmpc_b200.mpa import get_mpa
    from pdmpc_b200.records import SearchBatch
    mpa = get_mpa(args.mpa, non_convex=True)
    cores = os.cpu_count() or 1
    plan = lambda b: oracle_py.plan_batch(mpa, b, cores)
    batches = []
    for s in range(2):
        sc = scenario.commonroad_scenario(mpa, args.vehicles, seed=1 + s)
        batches.append(scenario.roll_out(sc, plan, args.sim_steps))
    base = SearchBatch.concat(batches)
    # bounded sample: replicate the two scenarios until one step is ~2 s of wall time
    t0 = time.perf_counter()
    oracle_py.plan_batch(mpa, base, cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    reps = int(min(max(1, round(2.0 / dt)), 256))
    sample = SearchBatch.concat([base] * reps)
    for _ in range(args.warmup):
        oracle_py.plan_batch(mpa, sample, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_py.plan_batch(mpa, sample, cores)
    el = time.perf_counter() - t0
    val = sample.n * args.steps / el
    desc = f"{sample.n} searches/step = {reps}x the records of 2 scenarios x {args.sim_steps} steps x {args.vehicles} vehicles"
    print(json.dumps({
        "impl": "reference", "metric": "vehicle-plans/sec", "value": val, "unit": "plans/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"CPM Lab road network, {args.vehicles} vehicles, coloring priorities, "
                               f"{args.mpa} MPA, Hp 6, InterX checker", "cpu_threads": cores},
        "cpu_baseline": {"value": val, "unit": "plans/s", "cores": cores, "kind": "port", "sample": desc,
                         "note": "C oracle (gcc -O2, no FMA contraction) restating the MATLAB reference; "
                                 "MATLAB R2023a itself is unavailable in this image"},
        "e2e": {"value": val, "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    from pdmpc_b200 import capi
    from pdmpc_b200.mpa import get_mpa
    from pdmpc_b200.records import BatchResult

    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    planner = capi.Planner(dev)        # raises if the CUDA library / device is missing
    mpa = get_mpa(args.mpa, non_convex=True)
    planner.upload_mpa(mpa)
    Hp = mpa.Hp

    t_gen = time.perf_counter()
    cache = os.path.join(ROOT, "build", f"bench_{args.mpa}_{args.vehicles}v_{args.scenarios}s_{args.sim_steps}t_seed"
                                        f"{1 + rank * args.scenarios}.npz")
    workers = args.gen_workers or max(1, min(12, (os.cpu_count() or 1) // max(world, 1)))
    batch, step_recs, ts_recs = build_records(dev, args.mpa, args.scenarios, 1 + rank * args.scenarios,
                                     args.vehicles, args.sim_steps, "" if args.no_cache else cache, workers)
    t_gen = time.perf_counter() - t_gen
    n = batch.n

    stream = torch.cuda.ExternalStream(planner.stream(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ------------------------------------
    planner.stage(batch)
    for _ in range(max(args.warmup, 3)):
        planner.run_staged()
    planner.sync()
    sampler = ClockSampler(dev)
    barrier()
    sampler.start()
    kernel_ms = []
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                     # L2 flush, outside the per-step events
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            planner.run_staged()
            e1.record(stream)
        e1.synchronize()
        kernel_ms.append(e0.elapsed_time(e1))
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    res = planner.fetch()
    stats = planner.stats()
    ms_step = float(np.mean(kernel_ms))
    t_local = torch.tensor([sum(kernel_ms)], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    total_ms = float(t_local.item())
    value = n * world * args.steps / (total_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ("e2e") ------------------
    import dataclasses
    pinned = []   # keep the pinned torch storages alive

    def pinned_like(a: np.ndarray) -> np.ndarray:
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True)
        pinned.append(t)
        return t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape)

    host_in = {}
    for f in dataclasses.fields(batch):
        a = getattr(batch, f.name)
        if isinstance(a, np.ndarray):
            host_in[f.name] = pinned_like(a)
            host_in[f.name][...] = a
    hb = dataclasses.replace(batch, **host_in)
    out = BatchResult.empty(n, Hp)
    for f in dataclasses.fields(out):
        a = getattr(out, f.name)
        if isinstance(a, np.ndarray):
            setattr(out, f.name, pinned_like(a))
    import ctypes as C
    bi, bo = capi.batch_in(hb), capi.batch_out(out)
    e2e_steps = max(3, min(args.steps, 10))

    def time_e2e(chunks: int, steps: int) -> float:
        planner.set_pipeline_chunks(chunks)
        for _ in range(2):
            planner._check(planner.lib.pdmpc_plan_batch(planner.h, C.byref(bi), C.byref(bo)))
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            planner._check(planner.lib.pdmpc_plan_batch(planner.h, C.byref(bi), C.byref(bo)))
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert np.array_equal(out.pop_hash, res.pop_hash), "e2e path and staged path disagree"
        return n * world * steps / float(t.item())

    e2e_serial = time_e2e(1, 3)                     # copy in, search, copy out, one after the other
    e2e_val = time_e2e(args.pipeline_chunks, e2e_steps)   # the library's default: chunked pipeline
    st2 = planner.stats()
    planner.set_pipeline_chunks(0)

    # ---- per-time-step latency (levels sequential, host buffers) -----------------
    # What the drop-in does for one 20-vehicle time step: one pdmpc_plan_batch call per
    # computation level (host buffers in, host buffers out), levels one after the other.
    # Replayed for the first (up to 8) scenarios of the rank, whose records come first.
    lat = []
    lat_levels = 0
    planner.set_cta_queue(True)   # what the MATLAB drop-in selects: valid-only queue in the CTA-per-search shape
    if rank == 0:
        by_step, off = {}, 0
        for step, _level, cnt in step_recs:
            lb = batch.select(np.arange(off, off + int(cnt)))
            ro = BatchResult.empty(lb.n, Hp)
            by_step.setdefault(int(step), []).append((lb, ro, capi.batch_in(lb), capi.batch_out(ro)))
            off += int(cnt)
        lat_levels = float(np.mean([len(v) for v in by_step.values()]))
        for rep in range(3):
            for k, levels in sorted(by_step.items()):
                t0 = time.perf_counter()
                for _lb, _ro, lbi, lbo in levels:
                    planner._check(planner.lib.pdmpc_plan_batch(planner.h, C.byref(lbi), C.byref(lbo)))
                if rep:
                    lat.append((time.perf_counter() - t0) * 1e3)

    # ---- the same time steps, ONE call each (pdmpc_plan_timestep) ------------------------------
    # Base iter_v of all 20 vehicles + predecessor lists + fallback areas in, all plans out: one
    # H2D, one dependency-ordered launch (searches wait for their own predecessors' flags), one D2H.
    lat1 = []
    if rank == 0 and ts_recs:
        calls = []
        for step, tb, td in ts_recs:
            ro = BatchResult.empty(tb.n, Hp)
            pidx = td.pred_idx if td.pred_idx.size else np.zeros(1, dtype=np.int32)
            dc = capi.TimestepDepsC(pred_ptr=capi._ptr(td.pred_ptr, capi._p_i32), pred_idx=capi._ptr(pidx, capi._p_i32),
                                    fb_npts=capi._ptr(td.fb_npts, capi._p_i32), fb_x=capi._ptr(td.fb_x, capi._p_f64),
                                    fb_y=capi._ptr(td.fb_y, capi._p_f64))
            calls.append((int(step), tb, td, pidx, ro, capi.batch_in(tb), dc, capi.batch_out(ro)))
        for rep in range(3):
            for _step, _tb, _td, _pidx, _ro, tbi, tdc, tbo in calls:
                t0 = time.perf_counter()
                planner._check(planner.lib.pdmpc_plan_timestep(planner.h, C.byref(tbi), C.byref(tdc), C.byref(tbo)))
                if rep:
                    lat1.append((time.perf_counter() - t0) * 1e3)
        # same searches, same answers as the level-by-level replay above (pop-order hashes per time step)
        for step, _tb, _td, _pidx, ro, *_ in calls:
            a = np.sort(np.concatenate([r.pop_hash for _lb, r, _i, _o in by_step[step]]))
            assert np.array_equal(a, np.sort(ro.pop_hash)), f"time step {step}: one-call path != level-by-level path"

    planner.set_cta_queue(False)

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------
    cpu = None
    if rank == 0 and world == 1:
        from oracle import oracle_py, parity
        cores = os.cpu_count() or 1
        m = args.cpu_sample or min(n, 14000)
        sample = batch.select(np.arange(m))
        t0 = time.perf_counter()
        ref = oracle_py.plan_batch(mpa, sample, cores)
        t_cpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        oracle_py.plan_batch(mpa, batch.select(np.arange(min(m, 2000))), 1)
        t_cpu1 = time.perf_counter() - t0
        # parity of the timed outputs against the checker, on the sample
        sub = dataclasses.replace(res, **{f.name: getattr(res, f.name)[:m] for f in dataclasses.fields(res)
                                          if isinstance(getattr(res, f.name), np.ndarray)})
        parity.compare(sub, ref)
        cpu = {"value": m / t_cpu, "unit": "plans/s", "cores": cores, "kind": "port",
               "sample": f"first {m} of the {n} timed search records, one pass, {cores} threads",
               "single_thread_plans_per_s": min(m, 2000) / t_cpu1,
               "parity_checked": True,
               "note": "C oracle restating the MATLAB reference (MATLAB R2023a unavailable here)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # DRAM traffic of the dominant kernel: one `ncu --set full` capture of this same launch
        # (same records, same shape), summarised under profiles/; null when the workload differs
        traffic = None
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json")))
            if int(cap.get("searches", -1)) == n and cap.get("mpa") == args.mpa:
                traffic = float(cap["dram_bytes_read"]) + float(cap["dram_bytes_write"])
        except Exception:
            pass
        fp64_peak = planner.measure_fp64_peak()
        alg_bytes = algorithmic_bytes(batch, stats, Hp)
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        f64 = fp64_ops(batch, stats, Hp)
        line = {
            "metric": "vehicle-plans/sec", "value": value, "unit": "plans/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"CPM Lab road network (BASELINE configs[1]), {args.vehicles} vehicles, "
                                   f"coloring priorities, {args.mpa} MPA, Hp {Hp}, InterX checker, "
                                   f"{args.scenarios} scenarios/GPU x {args.sim_steps} steps, pre-rolled closed loop",
                       "searches_per_step_per_gpu": n, "l2": "flushed between timed steps (256 MiB write)",
                       "record_generation_s": round(t_gen, 1)},
            "e2e": {"value": e2e_val, "unit": "plans/s", "h2d_bytes_per_step": int(st2.h2d_bytes),
                    "d2h_bytes_per_step": int(st2.d2h_bytes), "steps": e2e_steps,
                    "pipeline": "chunked copy/search overlap inside pdmpc_plan_batch (pdmpc_set_pipeline_chunks "
                                f"{args.pipeline_chunks}: 0 = library default)",
                    "value_without_pipeline": e2e_serial,
                    "h2d_ms": st2.h2d_ms, "kernel_ms": st2.kernel_ms, "d2h_ms": st2.d2h_ms},
            "gpu_launches": int(args.steps * 1),   # one persistent search kernel per timed step (staged path)
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "kernel": "pdmpc::search_kernel", "algorithmic_bytes_per_launch": alg_bytes,
                         "fp64": {"ops_per_launch": f64, "achieved_tops": f64 / (ms_step * 1e-3) / 1e12,
                                  "peak_tops_mul_add": fp64_peak[0], "peak_tflops_fma": fp64_peak[1],
                                  "frac": f64 / (ms_step * 1e-3) / 1e12 / fp64_peak[0],
                                  "peak_source": "measured in this run (pdmpc_measure_fp64_peak, register-only kernel)",
                                  "note": "mul/add/sqrt without FMA; latency-bound serial search, see DESIGN.md"}},
            "cpu_baseline": cpu,
            "search_stats": {"pops_per_plan": stats.total_pops / max(n, 1),
                             "nodes_per_plan": stats.total_nodes / max(n, 1),
                             "exhausted_frac": float(res.is_exhausted.mean()),
                             "obstacle_cols_per_pop": stats.total_obstacle_cols / max(stats.total_pops, 1)},
            "latency_ms_per_timestep": ({"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                                         "max": float(np.max(lat)), "n": len(lat), "levels_per_step": lat_levels,
                                         "what": "wall time of all computation levels of one 20-vehicle time step, one "
                                                 "pdmpc_plan_batch call per level, host buffers in and out"}
                                        if lat else None),
            "latency_ms_per_timestep_one_call": ({"p50": float(np.percentile(lat1, 50)), "p99": float(np.percentile(lat1, 99)),
                                                  "max": float(np.max(lat1)), "n": len(lat1),
                                                  "what": "wall time of ONE pdmpc_plan_timestep call per 20-vehicle time "
                                                          "step (predecessor hand-over on the device), host buffers in "
                                                          "and out; same time steps and answers as above"}
                                                 if lat1 else None),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    planner.close()


if __name__ == "__main__":
    main()
