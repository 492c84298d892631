#!/usr/bin/env python
"""bench.py — throughput of the B200 hot path (`Solver.process`, src/core/lib.zig:189-251) on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload pile100k|...]

One "step" = one process(dt = 1/60, sub_steps, iters) call over the whole world.  N = 1 runs configs[1] (`pile100k`,
the configuration the metric is quoted on); N > 1 runs one independent pile100k replica per GPU (a single world does
not shard — DESIGN.md "Multi-GPU") and, in the extra `batched` object, the sharded batched-worlds configuration
(configs[4]: 4,096 worlds of 256 bodies per GPU, weak scaling, no collective on the data path).  torch is plumbing
only: device selection, the stream, CUDA events, the barrier and the max-over-ranks reduction.

Prints ONE JSON line (rank 0).  `value` is device-timed with the state resident in HBM; `e2e` is the same metric
through the C ABI with HOST buffers (pinned), per step: H2D of per-body force/torque inputs, process(), D2H of the
body state.  `--impl reference` times the CPU restatement of the reference (oracle/, reference sweep order) on the
host cores instead — the Zig reference itself cannot be built in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_NOTE = "algorithmic bytes per SURVEY.md 8(d): unique compulsory traffic per launch, SoA, 4-byte scalars"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Libraries (NCCL, the CUDA runtime) may print to fd 1; the contract is ONE JSON line on stdout.  Everything except the
# final line is therefore sent to stderr: fd 1 is pointed at stderr for the whole run and the line goes to the saved fd.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# ---- workloads -------------------------------------------------------------------------------------------------------
def workload_table():
    from resolve2d_b200 import scenes
    return {
        "box1k": (scenes.build_box1k, 60),
        "pile100k": (scenes.build_pile100k, 200),
        "mixed1M": (scenes.build_mixed1M, 20),
        "pyramid20k": (scenes.build_pyramid20k, 30),
        "pile10k": (lambda s: scenes.build_pile(s, 200, 50), 60),
    }


# ---- algorithmic bytes --------------------------------------------------------------------------------------
def algorithmic_bytes(st, S, I):
    """Per process() call and kernel class, from the measured counts of the step."""
    N, T, E, P, M, K, C = st.n_bodies, st.n_buckets, st.n_entries, st.n_pairs, st.n_manifolds, st.n_points, st.n_joints
    M2 = max(K - M, 0)
    M1 = M - M2
    p = 3  # radix passes the SURVEY formula assumes for T <= 16M
    broad = 16 * N + 8 * E + 8 * E * (1 + 2 * p) + (4 * E + 4 * (T + 1)) + (8 * E + 4 * (T + 1) + 16 * N) + 8 * P
    narrow = 8 * P + 28 * N + 136 * M
    coloring = 52 * M1 + 76 * M2 + 16 * N            # pre-step + partition (the colouring rounds themselves are L2 work)
    integrate = S * (64 * N + 60 * N) + 16 * N
    contacts = S * I * (28 * M + 44 * K + 32 * N)
    joints = S * I * 44 * C
    return {"broadphase": broad, "narrowphase": narrow, "coloring": coloring, "integrate": integrate,
            "solve_contacts": contacts, "solve_joints": joints}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- clocks ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the CPU arm -----------------------------------------------------------------------------------------------------------
def cpu_steps_per_second(workload, steps, preroll):
    """Oracle in the reference's own sweep order, 1 thread (the reference has no threading)."""
    from oracle import ORDER_REFERENCE, OracleSolver
    from resolve2d_b200 import scenes
    build, _ = workload_table()[workload]
    s = OracleSolver(2.0, 4, order=ORDER_REFERENCE)
    cfg = build(s)
    S, I = cfg["sub_steps"], cfg["iters"]
    if preroll:
        s.timed_steps(scenes.DT, S, I, preroll)
    sec = s.timed_steps(scenes.DT, S, I, steps)
    return s.num_bodies(), steps / sec, sec, S, I


def cpu_batched_world_steps_per_second(n_worlds, steps, preroll):
    """cfg5 on the host: `n_worlds` independent oracle Solvers (reference sweep order), one world per task on all host
    cores (the reference is single-threaded per world; independent worlds are the only parallelism it has).
    ctypes releases the GIL inside orc_timed_steps, so plain threads run the C++ oracle in parallel."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ORDER_REFERENCE, OracleSolver
    from resolve2d_b200 import scenes
    cores = os.cpu_count() or 1
    worlds = []
    for w in range(n_worlds):
        s = OracleSolver(2.0, 4, order=ORDER_REFERENCE)
        scenes.build_batch_world(s, w)
        worlds.append(s)

    def run(s, k):
        return s.timed_steps(scenes.DT, 4, 4, k)
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda s: run(s, preroll), worlds))
        t0 = time.perf_counter()
        list(ex.map(lambda s: run(s, steps), worlds))
        sec = time.perf_counter() - t0
    return n_worlds * steps / sec, cores, sec


def run_reference(args, rank, world):
    if rank != 0:
        return
    build, preroll = workload_table()[args.workload]
    preroll = min(preroll, args.preroll if args.preroll is not None else preroll)
    t0 = time.time()
    n, sps, sec, S, I = cpu_steps_per_second(args.workload, args.steps, preroll + args.warmup)
    value = n * sps
    sample = (f"{args.steps} process() calls of {args.workload} after {preroll + args.warmup} untimed calls, "
              f"{sec:.1f} s of CPU work")
    line = {
        "impl": "reference", "metric": "body-steps/s", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / sps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "bodies": n, "sub_steps": S, "iters": I, "dt": 1 / 60,
                   "note": "C++ restatement of resolve2d's Solver (oracle/, reference sweep order), not the Zig binary: "
                           "no zig toolchain in the image"},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    emit(line)


# ---- our arm ------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from resolve2d_b200 import Batch, Solver, scenes

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def timed_steps(stepper, k):
        """k process() calls, L2 flushed before each, per-step CUDA events on the launching stream; returns seconds."""
        evs = []
        barrier()
        for _ in range(k):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            stepper()
            b.record(stream)
            evs.append((a, b))
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs) * 1e-3

    build, preroll = workload_table()[args.workload]
    if args.preroll is not None:
        preroll = args.preroll
    solver = Solver(2.0, 4, device=local_rank)
    solver.set_stream(stream.cuda_stream)
    cfg = build(solver)
    S, I = cfg["sub_steps"], cfg["iters"]
    n_bodies = solver.num_bodies()
    dt = scenes.DT
    for _ in range(preroll):            # scene formation (the pile), not part of warm-up or timing
        solver.process(dt, S, I)
    solver.reorder()                    # device memory order follows the formed pile (also redone every 1,024 calls)
    for _ in range(args.warmup):
        solver.process(dt, S, I)

    launches = [0]

    def step():
        solver.process(dt, S, I)
        launches[0] += solver.stats().n_launches

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStart()
    sec = max_over_ranks(timed_steps(step, args.steps))
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop() if sampler else None
    value = world * n_bodies * args.steps / sec
    st = solver.stats()

    # ---- e2e: HOST buffers in, HOST buffers out, every step, through the C ABI ----
    forces = torch.zeros((n_bodies, 3), dtype=torch.float32).pin_memory()
    out = {k: torch.empty(shape, dtype=torch.float32).pin_memory().numpy() for k, shape in
           (("pos", (n_bodies, 2)), ("angle", (n_bodies,)), ("momentum", (n_bodies, 2)), ("ang_momentum", (n_bodies,)))}
    out["id"] = None
    out["aabb"] = None
    forces_np = forces.numpy()
    k_e2e = max(3, min(args.steps, 50))
    for _ in range(3):
        solver.write_forces(forces_np)
        solver.process_read(dt, S, I, out)
    # three back-to-back segments of k_e2e steps each, host wall clock; the MEDIAN segment is reported (a host-side hiccup
    # — the loop is ~0.5 ms of Python, PCIe and GPU per step — otherwise moves the number by several per cent)
    segments = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            solver.write_forces(forces_np)          # H2D + scatter on a side stream, joined before the solver kernel
            solver.process_read(dt, S, I, out)      # the step, then the export of the new state behind it: one synchronisation
        torch.cuda.synchronize(dev)
        segments.append(max_over_ranks(time.perf_counter() - t0))
    e2e_sec = sorted(segments)[1]
    e2e_value = world * n_bodies * k_e2e / e2e_sec

    # ---- per-kernel-class device times over the same kind of steps (CUDA events around every launch) ----
    solver.profile_enable(True)
    k_prof = max(3, min(args.steps, 20))
    for _ in range(k_prof):
        flush_buf.zero_()
        solver.process(dt, S, I)
    prof = solver.profile_read(reset=True)
    solver.profile_enable(False)
    st = solver.stats()
    abytes = algorithmic_bytes(st, S, I)
    peak, peak_src = measured_peak()
    # the persistent cooperative solver runs the integrators and the joints inside the contact-sweep launch
    if prof.get("integrate", (0, 0))[1] == 0:
        abytes["solve_contacts"] += abytes["integrate"]
    if prof.get("solve_joints", (0, 0))[1] == 0:
        abytes["solve_contacts"] += abytes["solve_joints"]
    kernel_names = {"broadphase": "k_grid_cells / k_scan_chained / k_fine_pairs (+ k_list_buckets / k_sort_buckets / k_bucket_count / k_bucket_write when dynamic large bodies exist)",
                    "narrowphase": "k_narrow", "coloring": "k_color (sets the owner bitmaps) + k_scan_owners + k_partition_prestep",
                    "solve_contacts": "k_solve_tiles (single world without joints: tile-local momentum in shared memory) or "
                                      "k_solve_persistent; the substep loop: integrators + dataflow contact sweeps (+ joints)",
                    "integrate": "k_integrate_forces / k_integrate_positions", "solve_joints": "k_solve_joints"}
    kernels = {}
    for name, (ms, cnt) in prof.items():
        if cnt == 0:
            continue
        per_step_ms = ms / k_prof
        gbs = abytes[name] / (per_step_ms * 1e-3) / 1e9 if per_step_ms > 0 else 0.0
        kernels[name] = {"ms_per_step": per_step_ms, "launches_per_step": cnt / k_prof, "avg_launch_us": 1e3 * ms / cnt,
                         "algorithmic_bytes_per_step": abytes[name], "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    dk = kernels[dominant]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get(dominant)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": kernel_names.get(dominant, dominant), "kernel_class": dominant, "achieved": dk["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dk["achieved_gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                "per_launch": {"algorithmic_bytes": dk["algorithmic_bytes_per_step"] / dk["launches_per_step"],
                               "avg_launch_us": dk["avg_launch_us"]},
                "note": BYTES_NOTE}

    # ---- batched worlds (configs[4]), sharded: every rank owns its own block of worlds, no collective ----
    batched = None
    if args.batch_worlds > 0:
        nw = args.batch_worlds
        batch = Batch(nw, 2.0, 4, device=local_rank)
        batch.set_stream(stream.cuda_stream)
        first_world = rank * nw
        for w in range(nw):
            scenes.build_batch_world(batch.world(w), first_world + w)
        for _ in range(args.batch_preroll):
            batch.process(dt, 4, 4)
        batch.reorder()
        for _ in range(3):
            batch.process(dt, 4, 4)
        kb = max(3, min(args.steps, 30))
        bsec = max_over_ranks(timed_steps(lambda: batch.process(dt, 4, 4), kb))
        bst = batch.stats()
        batched = {"metric": "world-steps/s", "value": world * nw * kb / bsec, "unit": "world-steps/s",
                   "workload": f"batch{nw}x256 per GPU", "worlds_per_gpu": nw, "bodies_per_world": 256,
                   "n_gpus": world, "scaling": "weak", "steps": kb, "ms_per_step": 1e3 * bsec / kb,
                   "body_steps_per_s": world * batch.num_bodies() * kb / bsec, "colors": bst.n_colors,
                   "manifolds": bst.n_manifolds, "launches_per_step": bst.n_launches}
        batch.destroy()

    if batched is not None and rank == 0 and world == 1 and not args.no_cpu:
        n_cpu_worlds = 16 * (os.cpu_count() or 1)
        wsps, cores, csec = cpu_batched_world_steps_per_second(n_cpu_worlds, 100, args.batch_preroll)
        batched["cpu_baseline"] = {"value": wsps, "unit": "world-steps/s", "cores": cores, "kind": "port",
                                   "sample": f"{n_cpu_worlds} worlds x 100 process() calls after {args.batch_preroll} untimed calls, "
                                             f"one world per task on {cores} threads ({csec:.1f} s timed), oracle/ in the "
                                             "reference's sweep order"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        t0 = time.time()
        n, sps, csec, _, _ = cpu_steps_per_second(args.workload, args.cpu_steps, preroll)
        cpu = {"value": n * sps, "unit": "body-steps/s", "cores": 1, "kind": "port",
               "sample": f"{args.cpu_steps} process() calls of {args.workload} after {preroll} untimed calls "
                         f"({csec:.1f} s timed, {time.time() - t0:.1f} s total), oracle/ in the reference's sweep order",
               "ms_per_step": 1e3 / sps}

    if rank == 0:
        line = {
            "metric": "body-steps/s", "value": value, "unit": "body-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload if world == 1 else f"{args.workload} x {world} independent replicas",
                       "bodies": n_bodies, "sub_steps": S, "iters": I, "dt": 1 / 60, "preroll_steps": preroll,
                       "l2": "flushed before every timed step (256 MiB memset); per-step CUDA events on the launching stream",
                       "counts": {"N": st.n_bodies, "T": st.n_buckets, "E": st.n_entries, "P": st.n_pairs,
                                  "M": st.n_manifolds, "K": st.n_points, "colors": st.n_colors,
                                  "color_rounds": st.n_color_rounds, "joints": st.n_joints,
                                  "joint_colors": st.n_joint_colors}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "body-steps/s", "h2d_bytes_per_step": 12 * n_bodies,
                    "d2h_bytes_per_step": 24 * n_bodies, "steps": k_e2e, "ms_per_step": 1e3 * e2e_sec / k_e2e,
                    "segments_ms_per_step": [1e3 * t / k_e2e for t in segments], "aggregate": "median of 3 segments",
                    "path": "r2d_write_forces (pinned host) -> r2d_process_read (pinned host)"},
            "gpu_launches": launches[0],
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu,
            "batched": batched,
        }
        emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pile100k")
    ap.add_argument("--preroll", type=int, default=None, help="scene-formation steps before warm-up (default per workload)")
    ap.add_argument("--cpu-steps", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--batch-worlds", type=int, default=4096, help="worlds per GPU of the `batched` leg (0 = skip)")
    ap.add_argument("--batch-preroll", type=int, default=100)
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
