"""GPU tier (-m gpu): the CUDA path, called through the C ABI of libr2d_b200.so, against the CPU oracle on the same
seeded inputs.  Pass bars: candidate-pair sets bit-exact; manifold sets exact with contact geometry
bit-exact (tolerance 0 — same IEEE ops, no FMA, same trig); body state vs the oracle swept in the SAME colour order
bit-exact (stated tolerance: 0 ulp; anything else is a bug, not noise)."""
import numpy as np
import pytest

import hashing as H
from oracle import ORDER_COLORED, ORDER_REFERENCE, OracleSolver
from parity import assert_bodies_equal, assert_manifolds_equal, run_parity
from resolve2d_b200 import (MODE_FAST, MODE_REFERENCE_ORDER, OPT_SLEEP_CALLS, OPT_SLEEPING, OPT_WARM_START, Batch, R2DError,
                            ShardedBatch, Solver, scenes)

pytestmark = pytest.mark.gpu


def test_library_reports_device():
    import ctypes as C
    from resolve2d_b200 import _abi
    lib = _abi.load_library()
    n = C.c_int()
    assert lib.r2d_device_count(C.byref(n)) == 0 and n.value >= 1
    assert lib.r2d_abi_version() == 2


def test_gpu_0_3_many_boxes():
    cand, _ = run_parity(lambda: Solver(2.0, 4), scenes.setup_0_3_many_boxes, 240, check_every=20, what="0_3")
    assert cand.stats().n_launches > 0


def test_gpu_0_1_car_platformer():
    run_parity(lambda: Solver(2.0, 4), scenes.setup_0_1_car_platformer, 300, check_every=25, what="0_1")


def test_gpu_0_1_car_platformer_driven():
    """user torque / force / angular momentum through the setters every frame"""
    run_parity(lambda: Solver(2.0, 4), scenes.setup_0_1_car_platformer, 240, check_every=20, pre_step=scenes.drive_0_1,
               what="0_1 driven")


def test_gpu_golden_prefix_of_reference_binary():
    """Until the first real impacts the colour order cannot matter, so the CUDA path must reproduce the golden hashes
    of the reference's own binary (tests/golden/appendix_f.json) — here steps 1..40 of 0_3 (no joints).  (0_1 has coupled
    joints whose colour order differs from list order from the first step on, so it is compared with the oracle only.)"""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_f.json")))
    for name, setup, last in (("0_3_many_boxes", scenes.setup_0_3_many_boxes, 40),):
        s = Solver(2.0, 4)
        setup(s)
        by_step = {g["step"]: g for g in gold[name]["steps"]}
        for step in range(1, last + 1):
            s.process(scenes.DT, 4, 4)
            if step in by_step:
                g = by_step[step]
                b = s.read_bodies()
                assert f"{H.state_hash(b):016x}" == g["state"], (name, step)
                assert f"{H.aabb_hash(b):016x}" == g["aabb"], (name, step)
                assert f"{H.pairs_hash(s.read_pairs()):016x}" == g["pairs"], (name, step)
                st = s.stats()
                assert (st.n_pairs, st.n_manifolds, st.n_points) == (g["C"], g["M"], g["K"])
                assert st.n_entries <= g["E"]     # fine grid: small bodies list one home cell instead of their 4 m cells


def _golden_runs():
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wasm_golden.json")))
    return {k: v for k, v in gold.items() if not k.startswith("_")}


@pytest.mark.parametrize("name", sorted(_golden_runs()))
def test_gpu_reference_order_mode_reproduces_the_reference_binary(name):
    """R2D_MODE_REFERENCE_ORDER: the CUDA broadphase, narrowphase, integrators, joints and contact arithmetic with the
    reference's own SEQUENTIAL sweep order (manifolds in creation order, joints in list order).  Compared DIRECTLY with the
    outputs of the reference's shipped binary (tests/golden/wasm_golden.json, produced by executing resolve2d.wasm): state and
    AABB hashes after EVERY process() call of all six runs — 1,285 steps, both example scenes, all four joint kinds,
    exclusions, key-driven inputs, other (dt, sub_steps, iters), body removal — plus the raw body dumps, bit for bit.  No
    oracle and no colouring spec in between."""
    g = _golden_runs()[name]
    s = Solver(2.0, 4)
    s.set_mode(MODE_REFERENCE_ORDER)
    (scenes.setup_0_1_car_platformer if g["scene"] == "0_1" else scenes.setup_0_3_many_boxes)(s)
    dt = np.float32(1.0) / np.float32(g["dt_rate"])
    removals = {}
    for r in g["removals"]:
        _, step, ids = r.split(":")
        removals[int(step)] = [int(x) for x in ids.split(",")]
    for rec in g["steps"]:
        step = rec["step"]
        if step > 0:
            for rid in removals.get(step, []):
                s.remove_rigid_body(rid)
            if g["driven"]:
                scenes.drive_0_1(s)
            s.process(dt, g["sub_steps"], g["iters"])
        b = s.read_bodies()
        assert len(b["id"]) == rec["n"], f"{name} step {step}: body count"
        assert f"{H.state_hash(b):016x}" == rec["state"], f"{name} step {step}: state hash"
        assert f"{H.aabb_hash(b):016x}" == rec["aabb"], f"{name} step {step}: aabb hash"
        if "bodies" in rec:
            want = np.array(rec["bodies"], dtype=np.uint32)
            got = np.empty_like(want)
            got[:, 0] = b["id"]
            got[:, 1:3] = b["pos"].view(np.uint32)
            got[:, 3] = b["angle"].view(np.uint32)
            got[:, 4:6] = b["momentum"].view(np.uint32)
            got[:, 6] = b["ang_momentum"].view(np.uint32)
            got[:, 7:11] = b["aabb"].view(np.uint32)
            assert np.array_equal(got, want), f"{name} step {step}: raw body dump"


def test_gpu_box1k():
    """cfg1: 1,000 mixed bodies falling into a box; 600 steps = fall, pile, settle."""
    cand, orc = run_parity(lambda: Solver(2.0, 4), scenes.build_box1k, 600, check_every=50, what="box1k")
    assert cand.stats().n_manifolds > 1500


def test_gpu_dataflow_colouring_and_its_fallback():
    """single worlds are coloured by the dataflow kernel phase (no rounds); a body with more manifolds than its in-place
    list holds (a plank on 48 discs) chains the rest; one with more than a colouring word holds (120 discs) makes the
    same launch fall back to Jones-Plassmann rounds — same colours either way"""
    cand, _ = run_parity(lambda: Solver(2.0, 4), scenes.build_box1k, 60, check_every=10, what="box1k flow")
    assert cand.stats().n_color_rounds == 0
    cand, _ = run_parity(lambda: Solver(2.0, 4), scenes.build_hub, 90, check_every=10, what="hub")
    st = cand.stats()
    assert st.n_colors >= 40 and st.n_color_rounds == 0
    cand, _ = run_parity(lambda: Solver(2.0, 4), lambda s: scenes.build_hub(s, n_discs=120), 40, check_every=10, what="hub120")
    st = cand.stats()
    assert st.n_colors >= 100 and st.n_color_rounds > 0


def test_gpu_hub_body_beyond_256_colours_keeps_stepping():
    """300 manifolds on one non-static body (a plank on 300 discs) against the 256 colours of the boundary: the 44
    lowest-priority manifolds are dropped from the sweeps (colour R2D_COLOR_DROPPED, stats.n_dropped), everything else
    of the step proceeds — bit for bit like the oracle under the same rule."""
    def build(s):
        return scenes.build_hub(s, n_discs=300)
    cand, orc = run_parity(lambda: Solver(2.0, 4), build, 40, check_every=10, what="hub300")
    st = cand.stats()
    assert st.n_colors == 256 and st.n_dropped >= 40 and st.n_dropped == orc.stats().n_dropped
    assert np.count_nonzero(cand.read_manifolds()["color"] == 0xFFFFFFFD) == st.n_dropped


# ---- the reference's roadmap items (README.md:59-64), behind options, off by default ------------------------------------------
def test_gpu_warm_starting_matches_the_oracle_and_helps():
    """R2D_OPT_WARM_START: contacts persist across calls keyed by stable ids — the CUDA path (hash table on the device, warm
    terms looked up in the pre-step, first sweep of the call) against the oracle's std::unordered_map, bit for bit, on a
    settling box (contacts appear, flip reference faces, vanish), a pyramid with all four joint kinds, and with a body
    removed mid-run.  And it does what it is for: with ONE iteration per substep the warm-started pile sinks less."""
    opts = ((OPT_WARM_START, 1),)
    run_parity(lambda: Solver(2.0, 4), scenes.build_box1k, 160, check_every=20, what="box1k warm", options=opts)
    def build(s):
        return scenes.build_pyramid(s, base=24, n_spinners=3)
    run_parity(lambda: Solver(2.0, 4), build, 60, check_every=20, what="pyramid24 warm", options=opts)
    removed = []
    def remove(s):
        if len(removed) < 2 and s.num_bodies() == 1003 - len(removed) // 2 and s.stats().n_manifolds > 1500:
            s.remove_rigid_body(500)
            removed.append(1)
    cand, _ = run_parity(lambda: Solver(2.0, 4), scenes.build_box1k, 130, check_every=10, pre_step=remove, what="box1k warm, removal",
                         options=opts)
    assert cand.num_bodies() == 1002
    pen = {}
    for warm in (0, 1):
        s = Solver(2.0, 4)
        s.set_option(OPT_WARM_START, warm)
        scenes.build_box1k(s)
        for _ in range(400):
            s.process(scenes.DT, 4, 1)
        m = s.read_manifolds()
        pen[warm] = float(np.mean(-m["depth"][m["n_points"] > 0][:, 0]))
    print(f"mean penetration after 400 calls at 1 iteration: cold {pen[0]:.4f} m, warm {pen[1]:.4f} m")
    assert pen[1] < pen[0]


def test_gpu_sleeping_matches_the_oracle():
    """R2D_OPT_SLEEPING: bodies that stayed slow for 10 calls are static for the duration of a call, a fast body wakes what it
    touched (one hop per call), user forces wake — CUDA path vs the oracle bit for bit through falling, settling, sleeping
    and a kick that wakes part of the pile; then the same in a batch of worlds (k_world_broad / k_world_solve)."""
    opts = ((OPT_SLEEPING, 1), (OPT_SLEEP_CALLS, 10))
    kicked = []
    def kick(s):
        if s.stats().n_manifolds > 1900 and len(kicked) < 40:   # 20 calls of a sideways push on one body of the settled pile
            s.body_handle(400).set_force(4000.0, 1500.0)
            kicked.append(1)
    cand, orc = run_parity(lambda: Solver(2.0, 4), scenes.build_box1k, 420, check_every=30, pre_step=kick, what="box1k sleeping",
                           options=opts)
    assert kicked
    # a settled pile sleeps: the two resting bodies of a sleeping pair are both static for the call, so their pair is gone
    cold = Solver(2.0, 4)
    scenes.build_box1k(cold)
    for _ in range(420):
        cold.process(scenes.DT, 4, 4)
    for _ in range(800):
        cand.process(scenes.DT, 4, 4)
        cold.process(scenes.DT, 4, 4)
    print(f"candidate pairs after 1220 calls: {cand.stats().n_pairs} with sleeping, {cold.stats().n_pairs} without")
    assert cand.stats().n_pairs < 0.8 * cold.stats().n_pairs
    # joints: a sleeper is a static body for the call, and two-body joints write momentum into static bodies too (Q10)
    def build(s):
        return scenes.build_pyramid(s, base=16, n_spinners=2)
    run_parity(lambda: Solver(2.0, 4), build, 150, check_every=25, what="pyramid16 sleeping", options=opts)
    run_parity(lambda: Solver(2.0, 4), scenes.setup_0_1_car_platformer, 200, check_every=25, what="0_1 sleeping", options=opts)
    batch = Batch(96, 2.0, 4)
    batch.set_option(OPT_SLEEPING, 1)
    batch.set_option(OPT_SLEEP_CALLS, 10)
    oracles = {}
    for w in range(96):
        scenes.build_batch_world(batch.world(w), w, nx=10, ny=5)
        if w in (0, 50, 95):
            o = OracleSolver(2.0, 4, order=ORDER_COLORED)
            o.set_option(OPT_SLEEPING, 1)
            o.set_option(OPT_SLEEP_CALLS, 10)
            scenes.build_batch_world(o, w, nx=10, ny=5)
            oracles[w] = o
    for _ in range(200):
        batch.process(scenes.DT, 4, 4)
        for o in oracles.values():
            o.process(scenes.DT, 4, 4)
    for w, o in oracles.items():
        assert np.array_equal(batch.world(w).read_pairs(), o.read_pairs()), w
        assert_bodies_equal(batch.world(w).read_bodies(), o.read_bodies(), f"sleeping world {w}")


def test_gpu_solver_flavours_agree(monkeypatch):
    """pile_10k through the tile solver (default for a single world without joints), through the persistent dataflow
    sweep (R2D_TILE_SOLVER=0) and through the device-side decline of the tile solver (a tile with more record tasks
    than allowed -> the host re-launches the persistent sweep): all three bit-identical to the oracle"""
    build = lambda s: scenes.build_pile(s, 200, 50)
    run_parity(lambda: Solver(2.0, 4), build, 40, check_every=20, what="pile10k tiles")
    monkeypatch.setenv("R2D_TILE_SOLVER", "0")
    run_parity(lambda: Solver(2.0, 4), build, 40, check_every=20, what="pile10k persistent")
    monkeypatch.setenv("R2D_TILE_SOLVER", "1")
    monkeypatch.setenv("R2D_TILE_MAX_TASKS", "2")
    run_parity(lambda: Solver(2.0, 4), build, 40, check_every=20, what="pile10k tile solver declines")


def test_gpu_substep_iteration_variants():
    for S, I in ((1, 1), (1, 4), (2, 10), (4, 0)):
        run_parity(lambda: Solver(2.0, 4), scenes.build_box1k, 60, check_every=20, what=f"box1k S{S} I{I}", sub_steps=S,
                   iters=I)


def test_gpu_pyramid_with_joints():
    """cfg4 at reduced size (base row 40 -> 820 bodies, 4 joint rows, 6 spinners): all joint kinds, I = 10."""
    def build(s):
        return scenes.build_pyramid(s, base=40, n_spinners=6)
    cand, _ = run_parity(lambda: Solver(2.0, 4), build, 120, check_every=20, what="pyramid40")
    assert cand.stats().n_joints > 100 and cand.stats().n_joint_colors >= 2


def test_gpu_pile_10k():
    """cfg2 at 1/10 size (10,000 discs): pair sets / manifolds / state bit-exact through pile formation."""
    def build(s):
        return scenes.build_pile(s, 200, 50)
    run_parity(lambda: Solver(2.0, 4), build, 120, check_every=30, what="pile10k")


def test_gpu_mixed_20k_with_large_bodies():
    """cfg3 at reduced size: mixed discs/rects at any angle + large multi-cell rectangles (big-body grid walk)."""
    def build(s):
        return scenes.build_mixed(s, 200, 100, n_large=12)
    cand, _ = run_parity(lambda: Solver(2.0, 4), build, 60, check_every=15, what="mixed20k")
    assert cand.stats().n_colors >= 4


def test_gpu_bucket_broadphase_still_matches(monkeypatch):
    """R2D_BROADPHASE=buckets: every body through the hashed 4 m buckets (the original pipeline, still the path of
    batches with dynamic large bodies); the default is the fine grid, which every other test exercises."""
    monkeypatch.setenv("R2D_BROADPHASE", "buckets")
    def build(s):
        return scenes.build_mixed(s, 100, 40, n_large=6)
    cand, orc = run_parity(lambda: Solver(2.0, 4), build, 60, check_every=20, what="mixed4k buckets")
    assert cand.stats().n_entries == orc.stats().n_entries
    monkeypatch.delenv("R2D_BROADPHASE")
    cand, orc = run_parity(lambda: Solver(2.0, 4), build, 60, check_every=20, what="mixed4k fine")
    assert cand.stats().n_entries < orc.stats().n_entries


def test_gpu_fine_grid_large_large_and_small_large_pairs():
    """dynamic 8-16 m rectangles stacked in a column above a small mixed field: small-large pairs from step ~120,
    large-large pairs (bucket kernels on the large-only coarse buckets) from step ~180."""
    def build(s):
        return scenes.build_mixed(s, 40, 12, n_large=6)
    cand, _ = run_parity(lambda: Solver(2.0, 4), build, 260, check_every=20, what="mixed column")
    n = len(cand.read_bodies()["id"])
    pairs = np.asarray(cand.read_pairs()).reshape(-1, 2)
    big = pairs >= n - 6
    assert (big[:, 0] & big[:, 1]).any() and (big[:, 0] ^ big[:, 1]).any()


def test_gpu_fine_grid_far_from_the_origin():
    def build(s, shift):
        fac = s.entity_factory()
        fac.make_downwards_gravity(scenes.GRAVITY)
        rng = scenes.SplitMix64(7)
        d = scenes.descs_box(rng, 12, 6, origin=(shift - 8.0, 2.0))
        fac.make_bodies(np.concatenate([scenes._static_rect((shift, -1), 40, 2), d]))
        return {"sub_steps": 4, "iters": 4}
    for shift in (-70000.0, 65536.0 * 1.05 - 3.0, 3.0e6):
        run_parity(lambda: Solver(2.0, 4), lambda s: build(s, shift), 50, check_every=25, what=f"shift {shift}")


def test_gpu_batch_matches_standalone_worlds():
    """cfg5 at reduced size: 24 worlds stepped in one batch == each world stepped alone by the oracle."""
    n_worlds, steps = 24, 90
    batch = Batch(n_worlds, 2.0, 4)
    oracles = []
    for w in range(n_worlds):
        scenes.build_batch_world(batch.world(w), w)
        o = OracleSolver(2.0, 4, order=ORDER_COLORED)
        scenes.build_batch_world(o, w)
        oracles.append(o)
    for step in range(steps):
        batch.process(scenes.DT, 4, 4)
        for o in oracles:
            o.process(scenes.DT, 4, 4)
        if (step + 1) % 30 == 0:
            for w in (0, 7, n_worlds - 1):
                ws = batch.world(w)
                assert np.array_equal(ws.read_pairs(), oracles[w].read_pairs()), (step, w)
                assert_manifolds_equal(ws.read_manifolds(), oracles[w].read_manifolds(), f"batch world {w}")
    allb = batch.read_bodies()
    at = 0
    for w, o in enumerate(oracles):
        ob = o.read_bodies()
        n = len(ob["id"])
        got = {k: v[at:at + n] for k, v in allb.items()}
        assert_bodies_equal(got, ob, f"batch world {w}")
        at += n


@pytest.mark.parametrize("cache", [None, "0", "300", "device-wide broadphase"])
def test_gpu_large_batch_uses_world_solver_and_matches_oracle(monkeypatch, cache):
    """>= 74 worlds switches the substep loop to the CTA-per-world kernel (k_world_solve: one lane per contact point, the
    world's slots in shared memory); sampled worlds must still match the oracle bit for bit (the other batch test, with 24
    worlds, goes through the persistent dataflow kernel).  R2D_WORLD_CACHE=0 keeps every world's slots in its slice of the
    global record arrays, 300 only those of the worlds that have outgrown 300 slots — same bits every way."""
    if cache == "device-wide broadphase":     # default: k_world_broad, the grid of a world in shared memory
        monkeypatch.setenv("R2D_WORLD_BROAD", "0")
    elif cache is not None:
        monkeypatch.setenv("R2D_WORLD_CACHE", cache)
    n_worlds, steps = 96, 60
    batch = Batch(n_worlds, 2.0, 4)
    sample = (0, 1, 37, 95)
    oracles = {}
    for w in range(n_worlds):
        scenes.build_batch_world(batch.world(w), w)
        if w in sample:
            o = OracleSolver(2.0, 4, order=ORDER_COLORED)
            scenes.build_batch_world(o, w)
            oracles[w] = o
    for step in range(steps):
        batch.process(scenes.DT, 4, 4)
        for o in oracles.values():
            o.process(scenes.DT, 4, 4)
    for w, o in oracles.items():
        ws = batch.world(w)
        assert np.array_equal(ws.read_pairs(), o.read_pairs()), w
        assert_manifolds_equal(ws.read_manifolds(), o.read_manifolds(), f"world {w}")
        assert_bodies_equal(ws.read_bodies(), o.read_bodies(), f"world {w}")


def test_gpu_sharded_batch_equals_one_batch():
    """r2d_sharded_* (world w -> shard floor(w * G / n_worlds), one host thread + stream per shard, no
    collective): 200 worlds over 2 and over 3 shards — on two different GPUs when the box has them, else on one — stay
    bit-identical to the same 200 worlds in one batch, through process(), process_read() and write_forces()."""
    import ctypes as C
    from resolve2d_b200 import _abi
    n_dev = C.c_int()
    _abi.load_library().r2d_device_count(C.byref(n_dev))
    n_worlds = 200
    one = Batch(n_worlds, 2.0, 4)
    variants = [ShardedBatch(n_worlds, [0, 1 % n_dev.value]), ShardedBatch(n_worlds, [0, 1 % n_dev.value, 2 % n_dev.value])]
    assert variants[1].shard(1) == (67, 67) and variants[1].shard(2) == (134, 66)
    for w in range(n_worlds):
        d = scenes.descs_batch_world(w, nx=12, ny=6)
        for x in [one] + variants:
            fac = x.world(w).entity_factory()
            fac.make_downwards_gravity(scenes.GRAVITY)
            fac.make_bodies(d)
    n = one.num_bodies()
    rng = np.random.default_rng(11)
    for step in range(40):
        f = (rng.normal(size=(n, 3)) * 3).astype(np.float32)
        outs = []
        for x in [one] + variants:
            x.write_forces(f)
            if step % 3 == 0:
                outs.append(x.process_read(scenes.DT, 4, 4, x.read_bodies()))
            else:
                x.process(scenes.DT, 4, 4)
                outs.append(x.read_bodies())
        for k, o in enumerate(outs[1:]):
            assert_bodies_equal(o, outs[0], f"sharded variant {k} step {step}")
    assert variants[0].stats().n_manifolds == one.stats().n_manifolds > 5000
    assert np.array_equal(variants[1].world(150).read_pairs(), one.world(150).read_pairs())
    # page-locked destinations: every shard's k_world_solve writes its slice of the caller's arrays itself (from its own device)
    import torch
    pin = lambda *shape: torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
    for step in range(6):
        outs = []
        for x in [one] + variants:
            out = {"id": None, "pos": pin(n, 2), "angle": pin(n), "momentum": pin(n, 2), "ang_momentum": pin(n), "aabb": None}
            outs.append(x.process_read(scenes.DT, 4, 4, out))
        for k, o in enumerate(outs[1:]):
            for key in ("pos", "angle", "momentum", "ang_momentum"):
                assert np.array_equal(o[key].view(np.uint32), outs[0][key].view(np.uint32)), f"sharded variant {k}, pinned {key}, step {step}"
        ref = one.read_bodies()
        assert np.array_equal(outs[0]["pos"].view(np.uint32), ref["pos"].view(np.uint32))


def test_gpu_large_batch_with_joints_uses_per_world_colouring_and_dataflow_sweep():
    """>= 74 small worlds WITH joints: coloured per world (sorted greedy in shared memory), swept by the persistent
    dataflow kernel, whose per-body ranks come from the exported colour masks."""
    n_worlds, steps = 80, 40
    batch = Batch(n_worlds, 2.0, 4)
    sample = (0, 3, 41, 79)
    oracles = {}
    def build(s, w):
        return scenes.build_pyramid(s, base=6 + w % 3, n_spinners=1 + w % 2)
    for w in range(n_worlds):
        build(batch.world(w), w)
        if w in sample:
            o = OracleSolver(2.0, 4, order=ORDER_COLORED)
            build(o, w)
            oracles[w] = o
    for step in range(steps):
        batch.process(scenes.DT, 4, 10)
        for o in oracles.values():
            o.process(scenes.DT, 4, 10)
    for w, o in oracles.items():
        ws = batch.world(w)
        assert np.array_equal(ws.read_pairs(), o.read_pairs()), w
        assert_manifolds_equal(ws.read_manifolds(), o.read_manifolds(), f"world {w}")
        assert_bodies_equal(ws.read_bodies(), o.read_bodies(), f"world {w}")


def test_gpu_per_world_colouring_by_rounds_matches_sorted(monkeypatch):
    """R2D_WORLD_COLORING=rounds (Jones-Plassmann rounds per world) and the default (sort + sequential greedy) give the
    same colours, hence the same state."""
    def run():
        batch = Batch(96, 2.0, 4)
        for w in range(96):
            scenes.build_batch_world(batch.world(w), w)
        for _ in range(50):
            batch.process(scenes.DT, 4, 4)
        return batch.read_bodies(), batch.world(5).read_manifolds()
    b1, m1 = run()
    monkeypatch.setenv("R2D_WORLD_COLORING", "rounds")
    b2, m2 = run()
    assert_bodies_equal(b1, b2, "world colouring flavours")
    assert_manifolds_equal(m1, m2, "world colouring flavours")


def test_gpu_reorder_does_not_change_results():
    """The spatial device order is invisible: forcing a re-sort every step gives the same bits as never re-sorting."""
    a, b = Solver(2.0, 4), Solver(2.0, 4)
    a.set_reorder_interval(1)
    b.set_reorder_interval(0)
    for s in (a, b):
        scenes.build_box1k(s)
    for _ in range(40):
        a.process(scenes.DT, 4, 4)
        b.process(scenes.DT, 4, 4)
    assert_bodies_equal(a.read_bodies(), b.read_bodies(), "reorder")
    assert np.array_equal(a.read_pairs(), b.read_pairs())
    assert_manifolds_equal(a.read_manifolds(), b.read_manifolds(), "reorder")


def test_gpu_dataflow_colouring_list_flavour(monkeypatch):
    """Worlds with more candidate pairs than the register slots of the dataflow colouring hold (> 1.2 M) keep each
    thread's pending manifolds in a compacted list; R2D_FLOW_LIST=1 sends small worlds through that flavour: same
    colours, no rounds — hubs (chained lists) and mixed shapes included."""
    monkeypatch.setenv("R2D_FLOW_LIST", "1")
    for name, build, steps in (("hub", scenes.build_hub, 40), ("pile", lambda s: scenes.build_pile(s, 100, 40), 60),
                               ("mixed", lambda s: scenes.build_mixed(s, 80, 40, n_large=4), 40)):
        cand, _ = run_parity(lambda: Solver(2.0, 4), build, steps, check_every=10, what=f"{name}, list flavour")
        assert cand.stats().n_color_rounds == 0, name


def test_gpu_device_resort_equals_host_resort(monkeypatch):
    """The periodic re-sort runs on the device (keys, radix sort, permutation of the body arrays; the host only rebuilds
    the exclusion list and the joints).  R2D_RESORT_CHECK=1 makes every device re-sort compare its order with the one
    build_image derives from the same state (a difference fails the call), and the results must equal — bit for bit —
    those of the host re-sort (R2D_DEVICE_RESORT=0) and of never re-sorting: joints, exclusions, sleeping counters and a
    batch of several larger worlds ride along."""
    def single(build, interval, sleeping=False):
        s = Solver(2.0, 4)
        if sleeping:
            s.set_option(OPT_SLEEPING, 1)
        cfg = build(s) or {"sub_steps": 4, "iters": 4}
        s.set_reorder_interval(interval)
        for _ in range(45):
            s.process(scenes.DT, cfg["sub_steps"], cfg["iters"])
        s.reorder()
        s.process(scenes.DT, cfg["sub_steps"], cfg["iters"])
        out = s.read_bodies(), s.read_pairs(), s.read_manifolds()
        s.deinit()
        return out

    def batch(interval):
        b = Batch(3, 2.0, 4)
        for w in range(3):
            scenes.build_mixed(b.world(w), 30 + 5 * w, 20, n_large=2, seed=100 + w)
        b.set_reorder_interval(interval)
        for _ in range(30):
            b.process(scenes.DT, 4, 4)
        out = b.read_bodies(), b.world(2).read_pairs(), b.world(2).read_manifolds()
        b.destroy()
        return out

    cases = {
        "pyramid": lambda iv: single(lambda s: scenes.build_pyramid(s, base=24, n_spinners=3), iv),
        "car platformer": lambda iv: single(scenes.setup_0_1_car_platformer, iv),
        "box1k sleeping": lambda iv: single(scenes.build_box1k, iv, sleeping=True),
        "batch of 3 mixed worlds": batch,
    }
    for name, run in cases.items():
        monkeypatch.setenv("R2D_RESORT_CHECK", "1")
        monkeypatch.delenv("R2D_DEVICE_RESORT", raising=False)
        dev = run(5)
        never = run(0)
        monkeypatch.delenv("R2D_RESORT_CHECK")
        monkeypatch.setenv("R2D_DEVICE_RESORT", "0")
        host = run(5)
        for other, what in ((host, "host re-sort"), (never, "no re-sort")):
            assert_bodies_equal(dev[0], other[0], f"{name}: device re-sort vs {what}")
            assert np.array_equal(dev[1], other[1]), f"{name}: pairs, device re-sort vs {what}"
            assert_manifolds_equal(dev[2], other[2], f"{name}: device re-sort vs {what}")


def test_gpu_fast_mode_grid_parameters():
    """MODE_FAST honours cell_width / table_mult (the reference stores but ignores them, lib.zig:254-255); the oracle
    in the same mode must agree bit for bit, and the candidate set must equal the parity-mode one on this scene."""
    def mk():
        s = Solver(2.0, 4)
        s.set_mode(MODE_FAST)
        return s
    cand = mk()
    orc = OracleSolver(2.0, 4, order=ORDER_COLORED)
    orc.set_mode(MODE_FAST)
    ref = OracleSolver(2.0, 4, order=ORDER_COLORED)
    for s in (cand, orc, ref):
        scenes.build_box1k(s)
    for step in range(60):
        for s in (cand, orc, ref):
            s.process(scenes.DT, 4, 4)
    assert np.array_equal(cand.read_pairs(), orc.read_pairs())
    assert np.array_equal(cand.read_pairs(), ref.read_pairs())
    assert_bodies_equal(cand.read_bodies(), orc.read_bodies(), "fast mode")


def test_gpu_fast_mode_batch_of_small_worlds():
    """R2D_MODE_FAST with the per-world kernels (k_world_broad builds the grid of a world in shared memory with the caller's
    cell width and table multiplier): sampled worlds vs the oracle in the same mode."""
    n_worlds = 90
    batch = Batch(n_worlds, 3.0, 3)
    batch.set_mode(MODE_FAST)
    oracles = {}
    for w in range(n_worlds):
        scenes.build_batch_world(batch.world(w), w, nx=12, ny=6)
        if w in (0, 44, 89):
            o = OracleSolver(3.0, 3, order=ORDER_COLORED)
            o.set_mode(MODE_FAST)
            scenes.build_batch_world(o, w, nx=12, ny=6)
            oracles[w] = o
    for _ in range(80):
        batch.process(scenes.DT, 4, 4)
        for o in oracles.values():
            o.process(scenes.DT, 4, 4)
    assert batch.stats().n_launches == 4
    for w, o in oracles.items():
        ws = batch.world(w)
        assert np.array_equal(ws.read_pairs(), o.read_pairs()), w
        assert_manifolds_equal(ws.read_manifolds(), o.read_manifolds(), f"fast world {w}")
        assert_bodies_equal(ws.read_bodies(), o.read_bodies(), f"fast world {w}")


def test_gpu_fast_mode_small_cells_make_rectangles_large():
    """MODE_FAST with 1.3 m cells: discs (1.0 wide) stay in the fine grid, rectangles (diagonal 1.41) become LARGE and
    dynamic — half of the scene goes through the bucket kernels; the mode switch after an upload re-derives the fine cell."""
    cand, orc = Solver(1.3, 4), OracleSolver(1.3, 4, order=ORDER_COLORED)
    for s in (cand, orc):
        scenes.build_box1k(s)
        s.process(scenes.DT, 4, 4)
        s.set_mode(MODE_FAST)
    for step in range(60):
        cand.process(scenes.DT, 4, 4)
        orc.process(scenes.DT, 4, 4)
        if step % 20 == 19:
            assert np.array_equal(cand.read_pairs(), orc.read_pairs()), step
    assert_manifolds_equal(cand.read_manifolds(), orc.read_manifolds(), "fast 1.3")
    assert_bodies_equal(cand.read_bodies(), orc.read_bodies(), "fast 1.3")


def test_gpu_remove_body_swapremove_and_dangling_joint():
    s, o = Solver(2.0, 4), OracleSolver(2.0, 4, order=ORDER_COLORED)
    for x in (s, o):
        scenes.setup_0_1_car_platformer(x)
    for x in (s, o):
        for _ in range(30):
            x.process(scenes.DT, 4, 4)
        x.remove_rigid_body(50)     # a free box: the last body (id 110) moves into its slot
    assert s.body_id_at(50 - 0) == o.body_id_at(50)
    for x in (s, o):
        for _ in range(30):
            x.process(scenes.DT, 4, 4)
    assert_bodies_equal(s.read_bodies(), o.read_bodies(), "after swapRemove")
    with pytest.raises(R2DError) as e:
        s.remove_rigid_body(50)
    assert e.value.name == "NoSuchIdExists"
    s.remove_rigid_body(36)         # body of a FixedPosition joint -> InvalidRigidBodyId on the next process
    with pytest.raises(R2DError) as e:
        s.process(scenes.DT, 4, 4)
    assert e.value.name == "InvalidRigidBodyId"


def test_gpu_empty_and_tiny_worlds():
    s = Solver(2.0, 4)
    s.process(scenes.DT, 4, 4)                      # no bodies at all
    assert s.num_bodies() == 0 and len(s.read_pairs()) == 0
    fac = s.entity_factory()
    fac.make_downwards_gravity(9.82)
    from resolve2d_b200 import BodyOptions, DiscOptions
    h = fac.make_disc_body(BodyOptions(pos=(38, 19), mass=100, mu=0.5), DiscOptions(1.0))
    s.process(scenes.DT, 4, 4)
    st = h.get()                                    # the free-fall KAT of SURVEY F.2 (body 37, step 1)
    got = np.array([st.pos_y, st.momentum_y, st.aabb_y], np.float32).view(np.uint32).tolist()
    assert got == [0x4197fc82, 0xc182eeef, 0x4197fde8]
    s.clear()
    assert s.num_bodies() == 0
    h2 = s.entity_factory().make_disc_body(BodyOptions(pos=(0, 0), mass=1), DiscOptions(1.0))
    assert h2.id == 1                               # the id counter survives clear() (Q16)


def test_gpu_nan_or_infinite_pose_is_harmless():
    """A body whose pose becomes NaN / infinite never passes an AABB test again (as in the reference, which has no guard);
    it must not disturb anybody else — every other body still matches the oracle bit for bit — and nothing may read out of
    bounds or hang (fine grid: its home cell is clamped; buckets: its cell index saturates)."""
    for bad in (float("nan"), float("inf"), -float("inf")):
        cand, orc = Solver(2.0, 4), OracleSolver(2.0, 4, order=ORDER_COLORED)
        for s in (cand, orc):
            scenes.build_box1k(s)
            s.process(scenes.DT, 4, 4)
            s.body_handle(int(s.read_bodies()["id"][10])).set_pos(bad, 3.0)
        for _ in range(40):
            cand.process(scenes.DT, 4, 4)
            orc.process(scenes.DT, 4, 4)
        a, b = cand.read_bodies(), orc.read_bodies()
        keep = np.arange(len(a["id"])) != 10
        for k in ("pos", "angle", "momentum", "ang_momentum"):
            assert np.array_equal(a[k][keep].view(np.uint32), b[k][keep].view(np.uint32)), (bad, k)
        assert not np.isfinite(a["pos"][10, 0])


def test_gpu_ragged_batch_with_empty_and_single_body_worlds():
    """120 worlds of very different sizes in one batch — empty worlds, a single free body, a floor only, small and larger
    boxes — through the CTA-per-world kernels (colouring by sort, world solver): every non-empty world must match the
    oracle stepped alone, bit for bit."""
    from resolve2d_b200 import BodyOptions, DiscOptions
    n_worlds, steps = 120, 50
    batch = Batch(n_worlds, 2.0, 4)
    def build(s, w):
        kind = w % 6
        if kind == 0:
            return                                   # empty world
        fac = s.entity_factory()
        fac.make_downwards_gravity(scenes.GRAVITY)
        if kind == 1:
            fac.make_disc_body(BodyOptions(pos=(1.0 + w, 5.0), mass=3.0, mu=0.4), DiscOptions(0.7))   # free fall
        elif kind == 2:
            fac.make_bodies(scenes._static_rect((0, -1), 36, 2))                                      # a floor, nothing else
        else:
            scenes.build_batch_world(s, w, nx=3 + 4 * (kind - 3), ny=2 + 3 * (kind - 3))
    oracles = {}
    for w in range(n_worlds):
        build(batch.world(w), w)
        if w % 6 != 0 and w % 7 in (0, 3):
            o = OracleSolver(2.0, 4, order=ORDER_COLORED)
            build(o, w)
            oracles[w] = o
    for step in range(steps):
        batch.process(scenes.DT, 4, 4)
        for o in oracles.values():
            o.process(scenes.DT, 4, 4)
    assert len(oracles) >= 20
    for w, o in oracles.items():
        ws = batch.world(w)
        assert np.array_equal(ws.read_pairs(), o.read_pairs()), w
        assert_manifolds_equal(ws.read_manifolds(), o.read_manifolds(), f"ragged world {w}")
        assert_bodies_equal(ws.read_bodies(), o.read_bodies(), f"ragged world {w}")
    assert batch.world(0).num_bodies() == 0


def test_gpu_write_forces_and_bulk_read():
    s, o = Solver(2.0, 4), OracleSolver(2.0, 4, order=ORDER_COLORED)
    for x in (s, o):
        scenes.build_box1k(x)
    n = s.num_bodies()
    rng = np.random.default_rng(7)
    for step in range(20):
        f = rng.normal(size=(n, 3)).astype(np.float32) * 5
        for x in (s, o):
            x.write_forces(f)
            x.process(scenes.DT, 4, 4)
    assert_bodies_equal(s.read_bodies(), o.read_bodies(), "write_forces")


def test_gpu_buffers_grow_and_the_step_is_redone(monkeypatch):
    """R2D_TEST_SMALL_BUFFERS=1: the grid-entry and pair buffers start far too small; process() notices the overflow at its
    one synchronisation point, grows them and redoes the step (the broadphase does not modify body state and every later
    kernel does nothing on an overflowed attempt) — results must be unaffected, fine grid and bucket pipeline alike."""
    monkeypatch.setenv("R2D_TEST_SMALL_BUFFERS", "1")
    run_parity(lambda: Solver(2.0, 4), scenes.build_box1k, 60, check_every=20, what="box1k small buffers")
    def build(s):
        return scenes.build_mixed(s, 40, 12, n_large=6)
    run_parity(lambda: Solver(2.0, 4), build, 200, check_every=40, what="mixed small buffers")
    monkeypatch.setenv("R2D_BROADPHASE", "buckets")
    run_parity(lambda: Solver(2.0, 4), scenes.build_box1k, 40, check_every=20, what="box1k small buffers, buckets")


def test_gpu_process_read_equals_process_then_read():
    """r2d_process_read / r2d_batch_process_read (export enqueued behind the step, one synchronisation) == process() followed
    by read_bodies(), with pinned and with pageable destinations, through the tile solver and through its fallback."""
    import torch
    def run(build, steps, pinned, S=4, I=4):
        a, b = Solver(2.0, 4), Solver(2.0, 4)
        for s in (a, b):
            build(s)
        n = a.num_bodies()
        mk = (lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()) if pinned else \
             (lambda shape, dt: torch.empty(shape, dtype=dt).numpy())
        out = {"id": mk((n,), torch.int32).view(np.uint32), "pos": mk((n, 2), torch.float32), "angle": mk((n,), torch.float32),
               "momentum": mk((n, 2), torch.float32), "ang_momentum": mk((n,), torch.float32), "aabb": mk((n, 4), torch.float32)}
        for _ in range(steps):
            a.process_read(scenes.DT, S, I, out)
            b.process(scenes.DT, S, I)
        ref = b.read_bodies()
        for k in out:
            assert np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)), k
    run(scenes.setup_0_3_many_boxes, 40, True)
    run(scenes.setup_0_3_many_boxes, 40, False)
    run(lambda s: scenes.build_pile(s, 100, 30), 60, True)            # 3,000 bodies: tile solver
    run(lambda s: scenes.build_pyramid(s, base=16, n_spinners=2), 30, True, 4, 10)   # joints: persistent solver
    batch, batch2 = Batch(80, 2.0, 4), Batch(80, 2.0, 4)
    for w in range(80):
        scenes.build_batch_world(batch.world(w), w, nx=8, ny=4)
        scenes.build_batch_world(batch2.world(w), w, nx=8, ny=4)
    n = batch.num_bodies()
    out = {"id": None, "pos": torch.empty((n, 2), dtype=torch.float32).pin_memory().numpy(), "angle": np.empty(n, np.float32),
           "momentum": None, "ang_momentum": None, "aabb": None}
    for _ in range(30):
        batch.process_read(scenes.DT, 4, 4, out)
        batch2.process(scenes.DT, 4, 4)
    ref = batch2.read_bodies()
    assert np.array_equal(out["pos"].view(np.uint32), ref["pos"].view(np.uint32))
    assert np.array_equal(out["angle"].view(np.uint32), ref["angle"].view(np.uint32))
    # pos / angle / momentum / ang_momentum all page-locked: k_world_solve exports every world itself, in the caller's order
    # (the device order is re-sorted on the way: the export goes through host_of_dev)
    pin = lambda *shape: torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
    out = {"id": None, "pos": pin(n, 2), "angle": pin(n), "momentum": pin(n, 2), "ang_momentum": pin(n), "aabb": None}
    for step in range(30):
        if step % 7 == 3:
            batch.reorder()
        batch.process_read(scenes.DT, 4, 4, out)
        batch2.process(scenes.DT, 4, 4)
    ref = batch2.read_bodies()
    for k in ("pos", "angle", "momentum", "ang_momentum"):
        assert np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)), f"in-kernel export: {k}"


def _pile_metrics(s):
    b, m = s.read_bodies(), s.read_manifolds()
    dyn = b["pos"][3:, 1]                                   # (both scenes list their static bodies first)
    pen = float(-m["depth"][m["n_points"] > 0][:, 0].min()) if len(m) else 0.0
    return {"mean_y": float(b["pos"][:, 1].mean()), "min_y": float(dyn.min()), "max_penetration": pen, "manifolds": len(m),
            "p2": float(np.sum(b["momentum"].astype(np.float64) ** 2))}


@pytest.mark.parametrize("scene", ["0_3", "box1k"])
def test_gpu_colour_order_vs_reference_order_bounded_divergence(scene):
    """The graph-colour sweep order against the reference's list order (R2D_MODE_REFERENCE_ORDER, which reproduces the
    reference binary bit for bit): identical until real impacts start, chaotic body by body afterwards (metres) — Gauss-
    Seidel is order dependent — but the two runs must remain the same physics.  Asserted at steps 120 / 240 / 400: nothing
    tunnels through the floor, the piles have the same height (mean y within 5 % + 0.15 m), the same number of contacts
    (within 8 %), comparable residual motion and comparable worst penetration."""
    setup = scenes.setup_0_3_many_boxes if scene == "0_3" else scenes.build_box1k
    col, ref = Solver(2.0, 4), Solver(2.0, 4)
    ref.set_mode(MODE_REFERENCE_ORDER)
    for x in (col, ref):
        setup(x)
    first = None
    for step in range(1, 401):
        col.process(scenes.DT, 4, 4)
        ref.process(scenes.DT, 4, 4)
        d = float(np.max(np.abs(col.read_bodies()["pos"] - ref.read_bodies()["pos"]))) if step <= 130 or step % 40 == 0 else None
        if scene == "0_3" and step <= 40:
            assert d == 0.0, step                           # no impact yet: the order cannot matter
        if first is None and d is not None and d > 1e-3:
            first = step
        if step in (120, 240, 400):
            a, b = _pile_metrics(col), _pile_metrics(ref)
            print(f"{scene} step {step}: max |dpos| {d:.3g} m; colour order {a}; reference order {b}")
            for m in (a, b):
                assert m["min_y"] > 0.2, (step, m)                       # bodies rest ON the floor (top at y = 0)
                assert m["max_penetration"] < 0.7, (step, m)
            assert abs(a["mean_y"] - b["mean_y"]) <= 0.05 * max(a["mean_y"], b["mean_y"]) + 0.15, (step, a, b)
            assert abs(a["manifolds"] - b["manifolds"]) <= 0.08 * b["manifolds"], (step, a, b)
            assert max(a["p2"], b["p2"]) < 1000.0 or 1 / 3 < a["p2"] / b["p2"] < 3, (step, a, b)
            assert abs(a["max_penetration"] - b["max_penetration"]) < 0.25, (step, a, b)
    print(f"ordering effect on {scene}: max |dpos| first exceeds 1e-3 at step {first}")


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json's FULL sizes.  The scene is settled on the GPU, its state handed to the oracle, and the following
# process() calls are compared bit for bit (pairs, manifolds, colours, state); plus size-independent properties.
# ---------------------------------------------------------------------------------------------------------------------
def _full_size_handoff(build, settle_steps, check_steps, what):
    gpu = Solver(2.0, 4)
    cfg = build(gpu)
    S, I = cfg["sub_steps"], cfg["iters"]
    for _ in range(settle_steps):
        gpu.process(scenes.DT, S, I)
    orc = OracleSolver(2.0, 4, order=ORDER_COLORED)
    build(orc)
    state = gpu.read_bodies()
    assert np.array_equal(state["id"], orc.read_bodies()["id"])
    orc.load_state(state)
    for k in range(check_steps):
        gpu.process(scenes.DT, S, I)
        orc.process(scenes.DT, S, I)
        tag = f"{what} step {settle_steps + k + 1}"
        assert np.array_equal(gpu.read_pairs(), orc.read_pairs()), f"{tag}: candidate pairs"
        assert_manifolds_equal(gpu.read_manifolds(), orc.read_manifolds(), tag)
        assert_bodies_equal(gpu.read_bodies(), orc.read_bodies(), tag)
    return gpu


def test_gpu_full_size_pile100k():
    gpu = _full_size_handoff(scenes.build_pile100k, 180, 5, "pile100k")
    st = gpu.stats()
    assert st.n_bodies == 100003 and st.n_manifolds > 150000


def test_gpu_full_size_pyramid20k():
    gpu = _full_size_handoff(scenes.build_pyramid20k, 40, 5, "pyramid20k")
    st = gpu.stats()
    assert st.n_bodies == 19951 and st.n_joints > 1500


def test_gpu_full_size_mixed1M():
    """cfg3, settled (20000 x 50 lattice, 200 calls: the pile has formed and the large rectangles have landed on it): three
    consecutive process() calls bit for bit against the oracle."""
    gpu = _full_size_handoff(scenes.build_mixed1M, 200, 3, "mixed1M")
    st = gpu.stats()
    assert st.n_bodies == 1001003 and st.n_manifolds > 1500000


def test_gpu_full_size_batch4096_sampled_worlds_and_determinism():
    """cfg5 at full size: sampled worlds vs the oracle, and two identical batches stay bit-identical (no dependence
    on thread timing anywhere: atomics only ever feed order-independent results)."""
    n_worlds, steps = 4096, 70
    a, b = Batch(n_worlds, 2.0, 4), Batch(n_worlds, 2.0, 4)
    sample = tuple(range(0, 4096, 132)) + (4095,)   # 33 worlds
    oracles = {}
    for w in range(n_worlds):
        d = scenes.descs_batch_world(w)
        for x in (a, b):
            fac = x.world(w).entity_factory()
            fac.make_downwards_gravity(scenes.GRAVITY)
            fac.make_bodies(d)
        if w in sample:
            o = OracleSolver(2.0, 4, order=ORDER_COLORED)
            scenes.build_batch_world(o, w)
            oracles[w] = o
    for _ in range(steps):
        a.process(scenes.DT, 4, 4)
        b.process(scenes.DT, 4, 4)
        for o in oracles.values():
            o.process(scenes.DT, 4, 4)
    ba, bb = a.read_bodies(), b.read_bodies()
    assert_bodies_equal(ba, bb, "two identical batches")
    assert a.stats().n_manifolds == b.stats().n_manifolds > 500000
    for w, o in oracles.items():
        ws = a.world(w)
        assert np.array_equal(ws.read_pairs(), o.read_pairs()), w
        assert_manifolds_equal(ws.read_manifolds(), o.read_manifolds(), f"world {w}")
        assert_bodies_equal(ws.read_bodies(), o.read_bodies(), f"world {w}")
