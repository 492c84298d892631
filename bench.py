#!/usr/bin/env python
"""bench.py — throughput of the B200 hot path (`Solver.process`, src/core/lib.zig:189-251) on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload pile100k|...]

One "step" = one process(dt = 1/60, sub_steps, iters) call over the whole world / the whole batch of worlds.

N = 1   headline = configs[1] `pile100k` (the configuration BASELINE.json's metric is quoted on), body-steps/s; the same
        line carries `batched` (configs[4]: all 4,096 worlds of 256 bodies on this GPU, world-steps/s, with its own e2e,
        roofline and multi-core CPU baseline) and `configs` (configs[0], [2], [3]: box1k, mixed1M, pyramid20k — ms/step,
        body-steps/s, dominant kernel, 1-core CPU baseline).
N > 1   headline = configs[4] AS NAMED: 4,096 worlds in total, rank r owns worlds [r*4096/N, (r+1)*4096/N)
        (world w -> GPU floor(w*G/n_worlds)), no collective on the data path, `scaling: "strong"`,
        world-steps/s.  Extra keys keep the weak series (4,096 worlds per GPU, one pile100k replica per GPU) and
        `single_gpu_same_box`: all 4,096 worlds on rank 0's GPU alone, measured in the same run.
        A single large world does not shard (DESIGN.md "Multi-GPU": replicas only).

torch is plumbing only: device selection, the stream, CUDA events, the barrier and the max-over-ranks reduction.
`value` is device-timed with the state resident in HBM; `e2e` is the same metric through the C ABI with pinned HOST
buffers every step (H2D of per-body force/torque inputs, process(), D2H of the body state).  `--impl reference` times the
CPU restatement of the reference (oracle/, reference sweep order) on the host cores — the Zig reference itself cannot be
built in this image.  Prints ONE JSON line (rank 0).
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

N_WORLDS = 4096            # BASELINE.json configs[4]
BYTES_NOTE = ("algorithmic bytes = unique compulsory traffic per launch (SURVEY.md 8d), restated in DESIGN.md section 5 for the "
              "kernels that are actually built: every array a kernel class reads or writes is counted once per process() call")


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
    # name -> (builder, scene-formation steps before warm-up, CPU steps of the 1-core baseline)
    return {
        "box1k": (scenes.build_box1k, 120, 60),
        "pile100k": (scenes.build_pile100k, 200, 30),
        "mixed1M": (scenes.build_mixed1M, 200, 3),
        "pyramid20k": (scenes.build_pyramid20k, 30, 10),
        "pile10k": (lambda s: scenes.build_pile(s, 200, 50), 60, 30),
    }


KERNEL_NAMES = {
    "broadphase": "k_grid_cells<count> -> k_scan_chained -> k_grid_cells<fill> -> k_fine_pairs<count> -> k_scan_chained -> "
                  "k_fine_pairs<write> (+ bucket kernels when dynamic large bodies exist); batches of small worlds: "
                  "k_world_broad (one CTA per world, grid in shared memory)",
    "narrowphase": "k_narrow",
    "coloring": "k_color (dataflow greedy colouring) + k_scan_owners + k_partition_prestep; batches: k_color_worlds_seq",
    "solve_contacts": "the substep loop (integrators + contact sweeps + joints) in one launch: k_solve_tiles (single world "
                      "without joints), k_solve_persistent (joints / large worlds), k_world_solve (batches: one CTA per "
                      "world, slots in shared memory)",
    "integrate": "k_integrate_forces / k_integrate_positions (launch-per-colour A/B path only)",
    "solve_joints": "k_solve_joints (launch-per-colour A/B path only)",
}
# what a kernel class is bound by when it is NOT bandwidth (measured: DESIGN.md section 8); the roofline itself is HBM
LIMITER = {
    "solve_contacts": "latency: dependent per-contact chains and body hand-offs; working set resident in shared memory / L2",
    "coloring": "latency: dependent hand-offs of the greedy order",
    "broadphase": "latency below ~1M bodies (six dependent launches); bandwidth above",
    "narrowphase": "gather latency at 2 CTAs per SM (107 registers)",
}


# ---- algorithmic bytes ----------------------------------------------------------------------------
def algorithmic_bytes(st, S, I, n_rect, batch):
    """Unique compulsory bytes per process() call and kernel class, from the measured counts of the step.  Each array a
    class touches is counted once (inputs read once, outputs written once, scratch written once and read once)."""
    N, T, E, P, M, K, C = st.n_bodies, st.n_buckets, st.n_entries, st.n_pairs, st.n_manifolds, st.n_points, st.n_joints
    M2 = max(K - M, 0)
    if batch:
        # k_world_broad: in aabb, pos, shape; out pose, view (rects), ncells, pair_cnt, pairs
        broad = 48 * N + 24 * N + 64 * n_rect + 8 * P
        coloring = 4 * P + 24 * M + 4 * M                       # in m_color, m_hdr, m_prio; out m_color
        # k_world_solve: in pos, mom, frc, prop, shape (80 N), m_color + np (8 P), raw manifolds (m_hdr, g0, g1, r0: 64 M;
        # r1: 16 M2); out pos, mom, frc (48 N), aabb (16 N)
        solve = 80 * N + 8 * P + 64 * M + 16 * M2 + 64 * N
    else:
        # grid count/fill + 2 scans + fine pairs: in aabb, pos, shape (48 N); scratch pose, bkt, fcell (48 N, w + r), view
        # (64 per rect, w), ncells (4 N), bucket counts / starts (2 T each, 4 B, w + r), entries (24 E, w + r), pair counts
        # (4 N, w + r), parked partners (32 N, w + r); out pairs (8 P)
        broad = 48 * N + 2 * 48 * N + 64 * n_rect + 4 * N + 2 * 2 * (2 * T) * 4 + 2 * 24 * E + 2 * 4 * N + 2 * 32 * N + 8 * P
        # k_color: in m_hdr, m_prio, adjacency lists (8 per manifold end, w + r); out m_color, owner bits;
        # k_partition_prestep: in raw manifolds (88 M + 16 M2), prop (16 N); out records (104 M + 40 M2)
        coloring = 24 * M + 2 * 16 * M + 4 * M + (88 * M + 16 * M2 + 16 * N) + (104 * M + 40 * M2)
        # substep loop, load-once: records in (104 M + 40 M2), bodies in (pos, mom, frc, prop, shape 80 N) and out (64 N)
        solve = 104 * M + 40 * M2 + 80 * N + 64 * N + 48 * C
    # k_narrow: in pairs, shape, pose (gathered, once), view; out m_color (4 P) + raw manifold (88 M)
    narrow = 8 * P + 32 * N + 64 * n_rect + 4 * P + 88 * M
    streamed = S * (64 * N + 60 * N) + 16 * N + S * I * (28 * M + 44 * K + 32 * N) + S * I * 44 * C   # SURVEY 8d, every sweep from HBM
    return {"broadphase": broad, "narrowphase": narrow, "coloring": coloring, "solve_contacts": solve,
            "integrate": S * 124 * N + 16 * N, "solve_joints": S * I * 44 * C}, streamed


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(key):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tpath)).get(key, {})
    except Exception:
        return {}


# ---- clocks ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
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


# ---- the CPU legs (oracle/ = test infrastructure; here only as the reported baseline) ----------------------------------------
def cpu_world(workload, steps, preroll=0, state=None):
    """Oracle in the reference's own sweep order, 1 thread (the reference has no threading).  `state`: a read_bodies() dict
    of the formed scene handed over from the GPU run (saves the CPU pre-roll); else `preroll` calls are run first."""
    from oracle import ORDER_REFERENCE, OracleSolver
    from resolve2d_b200 import scenes
    build = workload_table()[workload][0]
    s = OracleSolver(2.0, 4, order=ORDER_REFERENCE)
    cfg = build(s)
    S, I = cfg["sub_steps"], cfg["iters"]
    if state is not None:
        s.load_state(state)
    elif preroll:
        s.timed_steps(scenes.DT, S, I, preroll)
    sec = s.timed_steps(scenes.DT, S, I, steps)
    return s.num_bodies(), steps / sec, sec, S, I


def cpu_batch(n_worlds, steps, preroll, first_world=0):
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
        scenes.build_batch_world(s, first_world + w)
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
    """The reference arm: same metric / config / unit as our arm at this N, on the host cores, rank 0 only."""
    if rank != 0:
        return
    t0 = time.time()
    cores = os.cpu_count() or 1
    n_gpus = max(world, args.gpus)
    if n_gpus > 1 or args.workload == "batch4096x256":
        n_cpu_worlds = min(N_WORLDS, 16 * cores)
        cpu_steps = max(args.steps, 20)
        wsps, cores, sec = cpu_batch(n_cpu_worlds, cpu_steps, args.batch_preroll + args.warmup)
        sample = (f"{n_cpu_worlds} of the {N_WORLDS} worlds x {cpu_steps} process() calls after {args.batch_preroll + args.warmup} "
                  f"untimed calls, one world per task on {cores} threads ({sec:.1f} s timed)")
        line = {
            "impl": "reference", "metric": "world-steps/s", "value": wsps, "unit": "world-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * N_WORLDS / wsps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": batch_config(n_gpus, N_WORLDS),
            "cpu_baseline": {"value": wsps, "unit": "world-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": wsps, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0,
            "note": "C++ restatement of resolve2d's Solver (oracle/, reference sweep order), not the Zig binary: no zig toolchain in the image",
        }
        emit(line)
        return
    _, preroll, _ = workload_table()[args.workload]
    if args.preroll is not None:
        preroll = args.preroll
    n, sps, sec, S, I = cpu_world(args.workload, args.steps, preroll + args.warmup)
    value = n * sps
    sample = (f"{args.steps} process() calls of {args.workload} after {preroll + args.warmup} untimed calls, "
              f"{sec:.1f} s of CPU work")
    line = {
        "impl": "reference", "metric": "body-steps/s", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / sps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": world_config(args.workload, n, S, I, preroll),
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
        "note": "C++ restatement of resolve2d's Solver (oracle/, reference sweep order), not the Zig binary: no zig toolchain in the image",
    }
    emit(line)


def world_config(workload, n_bodies, S, I, preroll):
    return {"workload": workload, "bodies": n_bodies, "sub_steps": S, "iters": I, "dt": 1 / 60, "preroll_steps": preroll}


def batch_config(n_gpus, n_worlds):
    return {"workload": f"batch{n_worlds}x256" + (f" sharded over {n_gpus} GPUs" if n_gpus > 1 else ""),
            "worlds": n_worlds, "bodies_per_world": 256, "worlds_per_gpu": n_worlds // max(n_gpus, 1), "sub_steps": 4, "iters": 4,
            "dt": 1 / 60, "parallelism": f"worlds sharded contiguously over {n_gpus} GPU(s), no collective on the data path"}


# ---- our arm ------------------------------------------------------------------------------------------------------------------
class Gpu:
    def __init__(self, local_rank, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.world = torch, dist, world
        torch.cuda.set_device(local_rank)
        self.index = local_rank
        self.dev = torch.device("cuda", local_rank)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def barrier(self, collective=True):
        if self.world > 1 and collective:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, x: float, op, collective=True) -> float:
        if self.world == 1 or not collective:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x, collective=True):
        return self.reduce(x, self.dist.ReduceOp.MAX, collective)

    profile_range = False   # --profile-range: cudaProfilerStart/Stop around the timed steps of the main leg only

    def timed_steps(self, stepper, k, collective=True):
        """k process() calls, L2 flushed before each, per-step CUDA events on the launching stream; returns seconds."""
        torch = self.torch
        evs = []
        self.barrier(collective)
        if self.profile_range:
            torch.cuda.cudart().cudaProfilerStart()
        for _ in range(k):
            self.flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            stepper()
            b.record(self.stream)
            evs.append((a, b))
        self.barrier(collective)
        if self.profile_range:
            torch.cuda.cudart().cudaProfilerStop()
            self.profile_range = False
        return sum(a.elapsed_time(b) for a, b in evs) * 1e-3

    def pinned(self, shape):
        return self.torch.empty(shape, dtype=self.torch.float32).pin_memory().numpy()


def measure(gpu: Gpu, obj, n_units, S, I, steps, warmup, n_rect, batch, traffic_key, collective=True, e2e=True):
    """Device-timed steps, end-to-end steps and the per-kernel-class profile of one Solver / Batch that is already formed
    (pre-rolled and re-sorted).  `n_units` = bodies (single world) or worlds (batch) this rank steps."""
    from resolve2d_b200 import scenes
    dt = scenes.DT
    n_bodies = obj.num_bodies()
    for _ in range(warmup):
        obj.process(dt, S, I)
    launches = [0]

    def step():
        obj.process(dt, S, I)
        launches[0] += obj.stats().n_launches
    sec = gpu.max(gpu.timed_steps(step, steps, collective), collective)
    out = {"sec": sec, "ms_per_step": 1e3 * sec / steps, "launches": launches[0], "n_bodies": n_bodies}
    # ---- e2e: HOST buffers in, HOST buffers out, every step, through the C ABI ----
    if e2e:
        forces = gpu.pinned((n_bodies, 3))
        forces[:] = 0
        host = {"pos": gpu.pinned((n_bodies, 2)), "angle": gpu.pinned((n_bodies,)), "momentum": gpu.pinned((n_bodies, 2)),
                "ang_momentum": gpu.pinned((n_bodies,)), "id": None, "aabb": None}
        k_e2e = max(3, min(steps, 50))
        for _ in range(3):
            obj.write_forces(forces)
            obj.process_read(dt, S, I, host)
        # three back-to-back segments of k_e2e steps each, host wall clock; the MEDIAN segment is reported (a host-side
        # hiccup otherwise moves the number by several per cent)
        segments = []
        for _ in range(3):
            gpu.barrier(collective)
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                obj.write_forces(forces)          # H2D + scatter on a side stream, joined before the solver kernel
                obj.process_read(dt, S, I, host)  # the step, then the export of the new state behind it
            gpu.torch.cuda.synchronize(gpu.dev)
            segments.append(gpu.max(time.perf_counter() - t0, collective))
        e2e_sec = sorted(segments)[1]
        out["e2e"] = {"sec_per_step": e2e_sec / k_e2e, "steps": k_e2e, "h2d_bytes_per_step": 12 * n_bodies,
                      "d2h_bytes_per_step": 24 * n_bodies, "ms_per_step": 1e3 * e2e_sec / k_e2e,
                      "segments_ms_per_step": [1e3 * t / k_e2e for t in segments], "aggregate": "median of 3 segments",
                      "path": "write_forces (pinned host) -> process_read (pinned host)"}
    # ---- per-kernel-class device times over the same kind of steps (CUDA events around every launch) ----
    obj.profile_enable(True)
    k_prof = max(3, min(steps, 20))
    for _ in range(k_prof):
        gpu.flush_buf.zero_()
        obj.process(dt, S, I)
    prof = obj.profile_read(reset=True)
    obj.profile_enable(False)
    st = obj.stats()
    abytes, streamed = algorithmic_bytes(st, S, I, n_rect, batch)
    peak, peak_src = measured_peak()
    traffic = load_traffic(traffic_key)
    kernels = {}
    for name, (ms, cnt) in prof.items():
        if cnt == 0:
            continue
        per_step_ms = ms / k_prof
        gbs = abytes[name] / (per_step_ms * 1e-3) / 1e9 if per_step_ms > 0 else 0.0
        kernels[name] = {"ms_per_step": per_step_ms, "launches_per_step": cnt / k_prof, "avg_launch_us": 1e3 * ms / cnt,
                         "algorithmic_bytes_per_step": abytes[name], "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
        if traffic.get(name):
            kernels[name]["dram_bytes_per_step"] = traffic[name]
            kernels[name]["dram_over_algorithmic"] = traffic[name] / abytes[name]
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    dk = kernels[dominant]
    out["kernels"] = kernels
    out["roofline"] = {"bound": "hbm", "kernel_class": dominant, "kernel": KERNEL_NAMES.get(dominant, dominant),
                       "achieved": dk["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dk["achieved_gbs"] / peak,
                       "traffic": traffic.get(dominant), "peak_source": peak_src,
                       "per_launch": {"algorithmic_bytes": dk["algorithmic_bytes_per_step"] / dk["launches_per_step"],
                                      "avg_launch_us": dk["avg_launch_us"]},
                       "limiter": LIMITER.get(dominant),
                       "note": BYTES_NOTE}
    if traffic.get(dominant):   # the same fraction on MEASURED DRAM bytes (one ncu --set full capture of the workload)
        out["roofline"]["dram"] = {"bytes": traffic[dominant], "gbs": traffic[dominant] / (dk["ms_per_step"] * 1e-3) / 1e9,
                                   "frac": traffic[dominant] / (dk["ms_per_step"] * 1e-3) / 1e9 / peak}
    if dominant == "solve_contacts":
        out["roofline"]["if_every_sweep_streamed_from_hbm"] = {"bytes": streamed + abytes["integrate"],
                                                               "gbs": (streamed + abytes["integrate"]) / (dk["ms_per_step"] * 1e-3) / 1e9}
    out["counts"] = {"N": st.n_bodies, "T": st.n_buckets, "E": st.n_entries, "P": st.n_pairs, "M": st.n_manifolds,
                     "K": st.n_points, "colors": st.n_colors, "color_rounds": st.n_color_rounds, "joints": st.n_joints,
                     "joint_colors": st.n_joint_colors, "launches_per_step": st.n_launches}
    return out


def make_world(gpu, workload, preroll):
    from resolve2d_b200 import Solver, scenes
    build = workload_table()[workload][0]
    solver = Solver(2.0, 4, device=gpu.index)
    solver.set_stream(gpu.stream.cuda_stream)
    cfg = build(solver)
    for _ in range(preroll):            # scene formation (the pile), not part of warm-up or timing
        solver.process(scenes.DT, cfg["sub_steps"], cfg["iters"])
    solver.reorder()                    # device memory order follows the formed pile (also redone every 1,024 calls)
    return solver, cfg


def make_batch(gpu, n_worlds, first_world, preroll):
    from resolve2d_b200 import Batch, scenes
    batch = Batch(n_worlds, 2.0, 4, device=gpu.index)
    batch.set_stream(gpu.stream.cuda_stream)
    for w in range(n_worlds):
        scenes.build_batch_world(batch.world(w), first_world + w)
    for _ in range(preroll):
        batch.process(scenes.DT, 4, 4)
    batch.reorder()
    return batch


BATCH_RECTS_PER_WORLD = 3 + 126   # 3 static box rects + the rectangles of the 23 x 11 mixed lattice ((ix + iy) odd)


def batch_leg(gpu, n_worlds, first_world, args, collective, e2e=True, total_worlds=None, steps=None):
    """cfg5 on this rank: `n_worlds` worlds starting at `first_world`; world-steps/s over all ranks that take part."""
    batch = make_batch(gpu, n_worlds, first_world, args.batch_preroll)
    kb = steps or max(3, min(args.steps, 30))
    m = measure(gpu, batch, n_worlds, 4, 4, kb, args.warmup, BATCH_RECTS_PER_WORLD * n_worlds, True, f"batch{n_worlds}x256",
                collective, e2e)
    total = total_worlds if total_worlds is not None else n_worlds
    leg = {"metric": "world-steps/s", "unit": "world-steps/s", "value": total * kb / m["sec"], "steps": kb,
           "ms_per_step": m["ms_per_step"], "worlds_on_this_gpu": n_worlds, "worlds_total": total,
           "body_steps_per_s": total * 256 * kb / m["sec"], "counts": m["counts"], "kernels": m["kernels"],
           "roofline": m["roofline"], "gpu_launches": m["launches"]}
    if e2e:
        leg["e2e"] = dict(m["e2e"], value=total / m["e2e"]["sec_per_step"], unit="world-steps/s")
        del leg["e2e"]["sec_per_step"]
    batch.destroy()
    return leg


def world_leg(gpu, workload, args, collective=False, e2e=True, cpu=True, steps=None, replicas=1):
    _, preroll, cpu_steps = workload_table()[workload]
    if args.preroll is not None:
        preroll = args.preroll
    solver, cfg = make_world(gpu, workload, preroll)
    S, I = cfg["sub_steps"], cfg["iters"]
    steps = steps or args.steps
    state = solver.read_bodies() if cpu else None
    m = measure(gpu, solver, solver.num_bodies(), S, I, steps, args.warmup, cfg.get("n_rect", 0), False, workload, collective, e2e)
    n = m["n_bodies"]
    leg = {"metric": "body-steps/s", "unit": "body-steps/s", "value": replicas * n * steps / m["sec"], "steps": steps,
           "ms_per_step": m["ms_per_step"], "config": world_config(workload, n, S, I, preroll), "counts": m["counts"],
           "kernels": m["kernels"], "roofline": m["roofline"], "gpu_launches": m["launches"]}
    if e2e:
        leg["e2e"] = dict(m["e2e"], value=replicas * n / m["e2e"]["sec_per_step"], unit="body-steps/s")
        del leg["e2e"]["sec_per_step"]
    solver.deinit()
    if cpu:
        t0 = time.time()
        # the formed scene is handed to the oracle (read_bodies -> load_state): no CPU pre-roll
        _, sps, csec, _, _ = cpu_world(workload, args.cpu_steps or cpu_steps, state=state)
        leg["cpu_baseline"] = {"value": n * sps, "unit": "body-steps/s", "cores": 1, "kind": "port", "ms_per_step": 1e3 / sps,
                               "sample": f"{args.cpu_steps or cpu_steps} process() calls of {workload} from the state the GPU run formed "
                                         f"after {preroll} calls ({csec:.1f} s timed, {time.time() - t0:.1f} s total), oracle/ in the "
                                         "reference's sweep order"}
    return leg


def compact(leg):
    """A short form of a leg for the `configs` object."""
    dom = leg["roofline"]["kernel_class"]
    out = {"ms_per_step": leg["ms_per_step"], "body_steps_per_s": leg["value"], "steps": leg["steps"], "config": leg["config"],
           "counts": leg["counts"], "dominant_kernel": dom, "dominant_kernel_ms": leg["kernels"][dom]["ms_per_step"],
           "kernels_ms": {k: round(v["ms_per_step"], 4) for k, v in leg["kernels"].items()},
           "frac_of_hbm_peak": {k: round(v["frac_of_hbm_peak"], 4) for k, v in leg["kernels"].items()}}
    if "e2e" in leg:
        out["e2e_ms_per_step"] = leg["e2e"]["ms_per_step"]
    if "cpu_baseline" in leg:
        out["cpu_baseline"] = leg["cpu_baseline"]
    return out


def run_ours(args, rank, world, local_rank):
    gpu = Gpu(local_rank, world)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    gpu.profile_range = args.profile_range

    if world == 1 and args.workload != "batch4096x256":
        # ---- N = 1: pile100k headline + batched + the other configs ----
        main = world_leg(gpu, args.workload, args, cpu=not args.no_cpu)
        clocks = sampler.stop() if sampler else None
        batched = None
        if args.batch_worlds > 0:
            batched = batch_leg(gpu, args.batch_worlds, 0, args, collective=False)
            batched["config"] = batch_config(1, args.batch_worlds)
            if not args.no_cpu:
                n_cpu_worlds = 16 * (os.cpu_count() or 1)
                wsps, cores, csec = cpu_batch(n_cpu_worlds, 100, args.batch_preroll)
                batched["cpu_baseline"] = {"value": wsps, "unit": "world-steps/s", "cores": cores, "kind": "port",
                                           "sample": f"{n_cpu_worlds} worlds x 100 process() calls after {args.batch_preroll} untimed "
                                                     f"calls, one world per task on {cores} threads ({csec:.1f} s timed), oracle/ in "
                                                     "the reference's sweep order"}
        configs = {}
        for name in ([] if args.no_configs else ["box1k", "mixed1M", "pyramid20k"]):
            if name == args.workload:
                continue
            t0 = time.time()
            configs[name] = compact(world_leg(gpu, name, args, cpu=not args.no_cpu, steps=min(args.steps, 20)))
            log(f"[bench] {name}: {configs[name]['ms_per_step']:.3f} ms/step ({time.time() - t0:.0f} s)")
        line = {
            "metric": "body-steps/s", "value": main["value"], "unit": "body-steps/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(main["config"], counts=main["counts"],
                           l2="flushed before every timed step (256 MiB memset); per-step CUDA events on the launching stream"),
            "clocks": clocks, "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "roofline": main["roofline"],
            "kernels": main["kernels"], "cpu_baseline": main.get("cpu_baseline"), "batched": batched, "configs": configs,
        }
        if batched is not None:   # the N = 1 point of the multi-GPU series (the lines at N > 1 report this metric as `value`)
            line["strong_scaling_series"] = {"metric": "world-steps/s", "workload": f"batch{args.batch_worlds}x256", "n_gpus": 1,
                                             "value": batched["value"], "e2e_value": batched["e2e"]["value"]}
        emit(line)
        return

    # ---- N > 1 (or --workload batch4096x256): configs[4] as named, strong scaling ----
    total = args.batch_worlds if args.batch_worlds > 0 else N_WORLDS
    lo, hi = rank * total // world, (rank + 1) * total // world
    main = batch_leg(gpu, hi - lo, lo, args, collective=True, total_worlds=total, steps=args.steps)
    clocks = sampler.stop() if sampler else None
    extras = {}
    if world > 1 and not args.no_extras:
        weak = batch_leg(gpu, total, rank * total, args, collective=True, e2e=False, total_worlds=total * world)
        extras["weak_4096_worlds_per_gpu"] = {k: weak[k] for k in ("value", "unit", "ms_per_step", "worlds_on_this_gpu", "worlds_total")}
        rep = world_leg(gpu, "pile100k", args, collective=True, e2e=False, cpu=False, steps=min(args.steps, 20), replicas=world)
        extras["pile100k_replicas"] = {"value": rep["value"], "unit": "body-steps/s", "ms_per_step": rep["ms_per_step"],
                                       "note": "one independent pile100k world per GPU (a single world does not shard)"}
        gpu.barrier()
        if rank == 0:   # the same 4,096 worlds on ONE GPU of this box, for the strong-scaling ratio (the other ranks idle)
            one = batch_leg(gpu, total, 0, args, collective=False, e2e=False)
            extras["single_gpu_same_box"] = {k: one[k] for k in ("value", "unit", "ms_per_step", "worlds_on_this_gpu")}
        gpu.barrier()
    if rank == 0:
        line = {
            "metric": "world-steps/s", "value": main["value"], "unit": "world-steps/s", "n_gpus": world, "steps": main["steps"],
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(batch_config(world, total), counts_rank0=main["counts"],
                           l2="flushed before every timed step (256 MiB memset); per-step CUDA events on the launching "
                              "stream, max over ranks"),
            "clocks": clocks, "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "roofline": main["roofline"],
            "kernels": main["kernels"], "cpu_baseline": None, "body_steps_per_s": main["body_steps_per_s"],
        }
        line.update(extras)
        line["strong_scaling_series"] = {"metric": "world-steps/s", "workload": f"batch{total}x256", "n_gpus": world,
                                         "value": main["value"], "e2e_value": main["e2e"]["value"],
                                         "n1_value_same_box": extras.get("single_gpu_same_box", {}).get("value")}
        emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pile100k")
    ap.add_argument("--preroll", type=int, default=None, help="scene-formation steps before warm-up (default per workload)")
    ap.add_argument("--cpu-steps", type=int, default=None, help="steps of the 1-core CPU baseline (default per workload)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the box1k / mixed1M / pyramid20k legs of the N = 1 line")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the weak series and the single-GPU comparison")
    ap.add_argument("--batch-worlds", type=int, default=N_WORLDS, help="worlds of the batched leg in total (0 = skip at N = 1)")
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
