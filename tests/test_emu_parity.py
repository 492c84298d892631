"""CPU-only tier: the per-thread kernel bodies (r2d_pipeline.cuh) and the host registry, driven serially by
tests/emu, must reproduce the oracle (coloured Gauss-Seidel order) bit for bit: candidate-pair sets, manifolds,
colours, joint order and body state."""
import numpy as np
import pytest

from emu import EmuBatch, EmuSolver
from oracle import ORDER_COLORED, OracleSolver
from parity import assert_bodies_equal, assert_manifolds_equal, run_parity
from resolve2d_b200 import R2DError, scenes


def test_emu_0_3_many_boxes():
    run_parity(lambda: EmuSolver(2.0, 4), scenes.setup_0_3_many_boxes, 120, check_every=10, what="0_3")


def test_emu_0_1_car_platformer_driven():
    run_parity(lambda: EmuSolver(2.0, 4), scenes.setup_0_1_car_platformer, 150, check_every=10,
               pre_step=scenes.drive_0_1, what="0_1 driven")


def test_emu_box1k():
    run_parity(lambda: EmuSolver(2.0, 4), scenes.build_box1k, 90, check_every=15, what="box1k")


def test_emu_colouring_by_rounds_only(monkeypatch):
    """R2D_EMU_FLOW=0: the Jones-Plassmann rounds alone (the path of worlds too large for the dataflow colouring)."""
    monkeypatch.setenv("R2D_EMU_FLOW", "0")
    cand, _ = run_parity(lambda: EmuSolver(2.0, 4), scenes.build_box1k, 40, check_every=10, what="box1k rounds")
    assert cand.stats().n_color_rounds > 0


def test_emu_dataflow_colouring_is_used_and_falls_back_on_hubs():
    cand, _ = run_parity(lambda: EmuSolver(2.0, 4), scenes.build_box1k, 40, check_every=10, what="box1k flow")
    assert cand.stats().n_color_rounds == 0          # coloured without rounds
    # a plank on 48 discs: more manifolds on one body than its in-place list holds -> the rest is chained, still no rounds
    cand, _ = run_parity(lambda: EmuSolver(2.0, 4), scenes.build_hub, 60, check_every=10, what="hub")
    st = cand.stats()
    assert st.n_colors >= 40 and st.n_color_rounds == 0
    # a plank on 120 discs needs more colours than a body's colouring word holds -> rounds, same colours as the oracle
    cand, _ = run_parity(lambda: EmuSolver(2.0, 4), lambda s: scenes.build_hub(s, n_discs=120), 30, check_every=10, what="hub120")
    st = cand.stats()
    assert st.n_colors >= 100 and st.n_color_rounds > 0


# ---- fine-grid broadphase (small bodies in a fine home-cell table, only large bodies in the hashed 4 m buckets) ----------
def test_emu_fine_grid_is_used_and_buckets_mode_still_matches(monkeypatch):
    cand, orc = run_parity(lambda: EmuSolver(2.0, 4), scenes.build_box1k, 30, check_every=10, what="box1k fine")
    assert cand.stats().n_entries < orc.stats().n_entries        # small bodies list one home cell, not their 4 m cells
    monkeypatch.setenv("R2D_EMU_BROADPHASE", "buckets")
    cand, orc = run_parity(lambda: EmuSolver(2.0, 4), scenes.build_box1k, 30, check_every=10, what="box1k buckets")
    assert cand.stats().n_entries == orc.stats().n_entries


def test_emu_fine_grid_with_dynamic_large_bodies():
    """mixed scene: small discs/rects at any angle + DYNAMIC 8-16 m rectangles (large-large pairs come from the bucket
    kernels, small-large pairs through the small body's own coarse buckets) falling onto the field."""
    def build(s):
        return scenes.build_mixed(s, 40, 12, n_large=6)
    # the large rectangles are stacked in one column above the field: the first lands at step ~120 (small-large
    # pairs), the next ones on top of it from step ~180 (large-large pairs)
    cand, orc = run_parity(lambda: EmuSolver(2.0, 4), build, 260, check_every=20, what="mixed fine")
    assert cand.stats().n_entries < orc.stats().n_entries
    n = len(cand.read_bodies()["id"])
    pairs = np.asarray(cand.read_pairs()).reshape(-1, 2)
    big = pairs >= n - 6
    assert (big[:, 0] & big[:, 1]).any() and (big[:, 0] ^ big[:, 1]).any()


def test_emu_fine_grid_pyramid_with_spinners_and_joints():
    def build(s):
        return scenes.build_pyramid(s, base=16, n_spinners=3)
    run_parity(lambda: EmuSolver(2.0, 4), build, 60, check_every=20, what="pyramid16 fine")


def test_emu_fine_grid_far_from_the_origin_and_negative_cells():
    """bodies at large negative / positive coordinates (cell indices wrap in the tag and in the bucket hash)."""
    def build(s, shift):
        fac = s.entity_factory()
        fac.make_downwards_gravity(scenes.GRAVITY)
        import numpy as np
        rng = scenes.SplitMix64(7)
        d = scenes.descs_box(rng, 12, 6, origin=(shift - 8.0, 2.0))
        floor = scenes._static_rect((shift, -1), 40, 2)
        fac.make_bodies(np.concatenate([floor, d]))
        return {"sub_steps": 4, "iters": 4}
    for shift in (-70000.0, 65536.0 * 1.05 - 3.0, 3.0e6):
        run_parity(lambda: EmuSolver(2.0, 4), lambda s: build(s, shift), 50, check_every=25, what=f"shift {shift}")


def test_emu_fine_grid_batch_matches_standalone_worlds():
    n_worlds = 5
    batch = EmuBatch(n_worlds, 2.0, 4)
    oracles = []
    for w in range(n_worlds):
        scenes.build_batch_world(batch.world(w), w, nx=9, ny=5)
        o = OracleSolver(2.0, 4, order=ORDER_COLORED)
        scenes.build_batch_world(o, w, nx=9, ny=5)
        oracles.append(o)
    for step in range(60):
        batch.process(scenes.DT, 4, 4)
        for o in oracles:
            o.process(scenes.DT, 4, 4)
    for w, o in enumerate(oracles):
        ws = batch.world(w)
        assert np.array_equal(ws.read_pairs(), o.read_pairs()), w
        assert_bodies_equal(ws.read_bodies(), o.read_bodies(), f"emu batch world {w}")


def test_emu_fine_grid_in_fast_mode_with_small_coarse_cells():
    """MODE_FAST honours cell_width: with 1.3 m coarse cells the discs (1.0 wide) are small and the rectangles (diagonal
    1.41) are LARGE and dynamic — half of the scene goes through the bucket kernels, half through the fine grid, and the
    mode switch after the upload re-derives the fine cell."""
    from resolve2d_b200 import MODE_FAST
    for cell in (2.0, 1.3):
        cand, orc = EmuSolver(cell, 4), OracleSolver(cell, 4, order=ORDER_COLORED)
        for s in (cand, orc):
            scenes.build_box1k(s)
            s.process(scenes.DT, 4, 4)      # one step in parity mode first
            s.set_mode(MODE_FAST)
        for step in range(45):
            cand.process(scenes.DT, 4, 4)
            orc.process(scenes.DT, 4, 4)
            if step % 15 == 14:
                assert np.array_equal(cand.read_pairs(), orc.read_pairs()), (cell, step)
        assert_bodies_equal(cand.read_bodies(), orc.read_bodies(), f"fast mode cell {cell}")


def test_emu_process_read_equals_process_then_read():
    a, b = EmuSolver(2.0, 4), EmuSolver(2.0, 4)
    for s in (a, b):
        scenes.setup_0_3_many_boxes(s)
    n = a.num_bodies()
    out = {"id": np.empty(n, np.uint32), "pos": np.empty((n, 2), np.float32), "angle": np.empty(n, np.float32),
           "momentum": np.empty((n, 2), np.float32), "ang_momentum": None, "aabb": np.empty((n, 4), np.float32)}
    for _ in range(30):
        a.process_read(scenes.DT, 4, 4, out)
        b.process(scenes.DT, 4, 4)
    ref = b.read_bodies()
    for k in ("id", "pos", "angle", "momentum", "aabb"):
        assert np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)), k


def test_emu_process_read_rejects_a_wrong_body_count():
    import ctypes as C
    s = EmuSolver(2.0, 4)
    scenes.setup_0_3_many_boxes(s)
    n = s.num_bodies()
    buf = np.empty((n, 2), np.float32)
    st = s._fn("process_read")(s._h, scenes.DT, 4, 4, None, C.c_void_p(buf.ctypes.data), None, None, None, None, n - 1)
    assert st != 0                                   # R2D_ERR_INVALID_ARGUMENT: n must equal the number of bodies
    st = s._fn("process_read")(s._h, scenes.DT, 4, 4, None, C.c_void_p(buf.ctypes.data), None, None, None, None, n)
    assert st == 0


def test_emu_nan_or_infinite_pose_is_harmless():
    """A body whose pose becomes NaN / infinite (the reference has no guard either: it simply never passes an AABB test
    again) must not disturb anybody else, with the fine grid and with the bucket pipeline: every other body still matches
    the oracle bit for bit, and nothing reads out of bounds or hangs."""
    for bad in (float("nan"), float("inf"), -float("inf")):
        cand, orc = EmuSolver(2.0, 4), OracleSolver(2.0, 4, order=ORDER_COLORED)
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


def test_sharded_batch_matches_one_oracle_per_world():
    """r2d_sharded_* (SURVEY 8e): 7 worlds over 3 shards — world w on shard floor(w * 3 / 7), one host thread per shard —
    step exactly like 7 standalone oracle Solvers; bulk arrays are world-major over all shards."""
    from emu import EmuShardedBatch
    n_worlds, n_shards = 7, 3
    sb = EmuShardedBatch(n_worlds, n_shards)
    assert [sb.shard(k) for k in range(n_shards)] == [(0, 3), (3, 2), (5, 2)]
    for w in range(n_worlds):
        assert w * n_shards // n_worlds == [k for k in range(n_shards) if sb.shard(k)[0] <= w < sum(sb.shard(k))][0]
    oracles = []
    for w in range(n_worlds):
        scenes.build_batch_world(sb.world(w), w, nx=6 + w % 3, ny=3)
        o = OracleSolver(2.0, 4, order=ORDER_COLORED)
        scenes.build_batch_world(o, w, nx=6 + w % 3, ny=3)
        oracles.append(o)
    n = sb.num_bodies()
    assert n == sum(o.num_bodies() for o in oracles)
    rng = np.random.default_rng(3)
    out = None
    for step in range(25):
        f = (rng.normal(size=(n, 3)) * 2).astype(np.float32)
        sb.write_forces(f)
        at = 0
        for o in oracles:
            k = o.num_bodies()
            o.write_forces(f[at:at + k])
            o.process(scenes.DT, 2, 3)
            at += k
        if step % 2:
            sb.process(scenes.DT, 2, 3)
            out = sb.read_bodies()
        else:
            out = sb.process_read(scenes.DT, 2, 3, sb.read_bodies())
    at = 0
    for w, o in enumerate(oracles):
        ob = o.read_bodies()
        k = len(ob["id"])
        assert_bodies_equal({key: v[at:at + k] for key, v in out.items()}, ob, f"sharded world {w}")
        assert_bodies_equal(sb.world(w).read_bodies(), ob, f"sharded world {w} via its handle")
        at += k
    assert sb.stats().n_bodies == n
    with pytest.raises(R2DError):
        EmuShardedBatch(2, 3)   # fewer worlds than shards


def test_emu_hub_body_beyond_256_colours_keeps_stepping():
    """A plank resting on 300 discs has 300 manifolds on ONE non-static body; the boundary has 256 colours (R2D_MAX_COLORS,
    a declared hard limit).  The 44 lowest-priority manifolds are reported with colour R2D_COLOR_DROPPED and left out of
    the sweeps; the step itself — every other contact, integration — goes on, bit for bit like the oracle under the same
    rule (round 1 returned ColorOverflow here and the world froze for good)."""
    def build(s):
        return scenes.build_hub(s, n_discs=300)
    cand, orc = run_parity(lambda: EmuSolver(2.0, 4), build, 12, check_every=4, what="hub300")
    st = cand.stats()
    assert st.n_colors == 256 and st.n_dropped >= 40 and st.n_dropped == orc.stats().n_dropped
    man = cand.read_manifolds()
    assert np.count_nonzero(man["color"] == 0xFFFFFFFD) == st.n_dropped


def test_emu_backend_resort_keeps_results_and_derives_the_host_order(monkeypatch):
    """BatchBase::reorder with a backend that re-sorts its own arrays (the CUDA backend does it on the device; the emulator
    restates it serially and always compares the order with build_image's): re-sorting every few steps must not change a
    bit — joints, exclusion pairs, a batch of several worlds — against the host path and against never re-sorting."""
    def single(build, interval):
        s = EmuSolver(2.0, 4)
        build(s)
        s.set_reorder_interval(interval)
        for k in range(40):
            if interval and k % 9 == 4:
                s.reorder()
            s.process(scenes.DT, 4, 4)
        return s.read_bodies(), s.read_pairs(), s.read_manifolds()

    def batch(interval):
        b = EmuBatch(3, 2.0, 4)
        for w in range(3):
            scenes.build_mixed(b.world(w), 12 + 2 * w, 8, n_large=1, seed=50 + w)
        b.set_reorder_interval(interval)
        for k in range(25):
            if interval and k % 9 == 4:
                b.reorder()
            b.process(scenes.DT, 4, 4)
        bodies = [b.world(w).read_bodies() for w in range(3)]
        merged = {k: np.concatenate([x[k] for x in bodies]) for k in bodies[0]}
        return merged, b.world(1).read_pairs(), b.world(1).read_manifolds()

    monkeypatch.setenv("R2D_EMU_RESORT", "2")      # the backend path is the one reorder() takes
    with pytest.raises(R2DError):
        single(scenes.setup_0_3_many_boxes, 3)
    for name, run in (("0_1", lambda iv: single(scenes.setup_0_1_car_platformer, iv)),
                      ("pyramid", lambda iv: single(lambda s: scenes.build_pyramid(s, base=10, n_spinners=2), iv)),
                      ("batch", batch)):
        monkeypatch.delenv("R2D_EMU_RESORT", raising=False)
        own = run(3)
        never = run(0)
        monkeypatch.setenv("R2D_EMU_RESORT", "0")
        host = run(3)
        for other, what in ((host, "host re-sort"), (never, "no re-sort")):
            assert_bodies_equal(own[0], other[0], f"{name}: backend re-sort vs {what}")
            assert np.array_equal(own[1], other[1]), f"{name}: pairs vs {what}"
            assert_manifolds_equal(own[2], other[2], f"{name}: manifolds vs {what}")
