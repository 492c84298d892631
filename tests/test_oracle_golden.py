"""Pins the CPU oracle against golden vectors taken from the reference's own prebuilt binary (SURVEY.md Appendix F;
tests/golden/appendix_f.json): per-step FNV hashes of body state and AABBs, candidate-pair sets, manifold lists and
raw f32 bit patterns for the reference's example scenes 0_1_car_platformer (plain and driven) and 0_3_many_boxes.
Everything is compared bit for bit."""
import json
import os

import numpy as np
import pytest

import hashing as H
from oracle import ORDER_REFERENCE, OracleSolver
from resolve2d_b200 import scenes

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_f.json")))
SETUP = {
    "0_1_car_platformer": (scenes.setup_0_1_car_platformer, None),
    "0_1_car_platformer_driven": (scenes.setup_0_1_car_platformer, scenes.drive_0_1),
    "0_3_many_boxes": (scenes.setup_0_3_many_boxes, None),
}


def _hex(v):
    return f"{v:016x}"


@pytest.mark.parametrize("name", list(SETUP))
def test_oracle_matches_reference_binary(name):
    setup, drive = SETUP[name]
    gold = GOLD[name]
    s = OracleSolver(2.0, 4, order=ORDER_REFERENCE)
    setup(s)
    b = s.read_bodies()
    assert _hex(H.state_hash(b)) == gold["step0"]["state"]
    assert _hex(H.aabb_hash(b)) == gold["step0"]["aabb"]
    by_step = {g["step"]: g for g in gold["steps"]}
    raw = {}
    for r in gold["raw"]:
        raw.setdefault(r["step"], []).append(r)
    for step in range(1, max(by_step) + 1):
        if drive:
            drive(s)
        s.process(scenes.DT, 4, 4)
        if step in by_step:
            g = by_step[step]
            b = s.read_bodies()
            st = s.stats()
            assert _hex(H.state_hash(b)) == g["state"], f"{name} step {step}: state"
            assert _hex(H.aabb_hash(b)) == g["aabb"], f"{name} step {step}: aabb"
            assert st.n_entries == g["E"], f"{name} step {step}: E"
            assert st.n_pairs == g["C"] and _hex(H.pairs_hash(s.read_pairs())) == g["pairs"], f"{name} step {step}: pairs"
            man = s.read_manifolds()
            assert (len(man), int(man["n_points"].sum())) == (g["M"], g["K"]), f"{name} step {step}: M,K"
            assert _hex(H.manifolds_hash(man)) == g["man"], f"{name} step {step}: manifolds"
        for r in raw.get(step, []):
            st = s.body_handle(r["id"]).get()
            got = np.array([st.pos_x, st.pos_y, st.angle, st.momentum_x, st.momentum_y, st.ang_momentum, st.aabb_x,
                            st.aabb_y, st.aabb_half_w, st.aabb_half_h], np.float32).view(np.uint32)
            want = [int(x, 16) for x in r["pos"] + [r["angle"]] + r["momentum"] + [r["ang_momentum"]] + r["aabb"]]
            assert got.tolist() == want, f"{name} step {step} body {r['id']}"


def test_reference_unit_tests_aabb(oracle_lib):
    """src/core/aabb.zig:34-45 'sanity check aabb overlapping' — the only KAT the reference ships for this path."""
    f = oracle_lib.orc_aabb_intersects
    assert f(1, 1, 1.0, 0.5, 2, 3, 1.5, 3.0) == 1
    assert f(1, 1, 1.0, 0.5, 2, 10, 1.5, 3.0) == 0
    assert f(1, 10, 1.0, 0.5, 2, 10, 1.5, 3.0) == 1


def test_trig_known_answers(oracle_lib):
    """Small-angle early-outs and quadrant reduction of the compiler-rt sinf/cosf."""
    s, c = oracle_lib.orc_sinf, oracle_lib.orc_cosf
    assert s(0.0) == 0.0 and c(0.0) == 1.0
    tiny = float(np.float32(2.0 ** -13))
    assert s(tiny) == tiny and c(tiny) == 1.0
    xs = np.concatenate([np.linspace(-20, 20, 4001), [0.1, 1.0, 3.0, -0.3, 16.0]]).astype(np.float32)
    got_s = np.array([s(float(x)) for x in xs], np.float32)
    got_c = np.array([c(float(x)) for x in xs], np.float32)
    assert np.max(np.abs(got_s - np.sin(xs.astype(np.float64)))) < 1.2e-7
    assert np.max(np.abs(got_c - np.cos(xs.astype(np.float64)))) < 1.2e-7
    # golden: rock at angle 1.0 / spinner aabb come from the wasm run and are covered by the scene test above


def test_cell_hash(oracle_lib):
    """SpatialHash.hash (SpatialHash.zig:78-81) incl. negative cells: u64 bitcast of i64 products, `*` before `^`."""
    h = oracle_lib.orc_cell_hash
    for xi, yi, t in [(0, 0, 1046), (3, -2, 1046), (-7, 11, 222), (-1, -1, 2_000_000), (123456, -98765, 200006)]:
        want = ((xi * 92837111) ^ (yi * 689287499)) % (1 << 64) % t
        assert h(t, xi, yi) == want


def test_magic_modulo_equals_the_64_bit_modulo():
    """cell_hash_magic (mulhi by floor((2^64 - 1) / T) + two corrections) == u64(xi * 92837111 ^ yi * 689287499) % T, the
    reference's hash (SpatialHash.zig:78-81): checked against Python's big integers on random and extreme operands."""
    import random
    rng = random.Random(5)
    M64 = (1 << 64) - 1
    cases = [(0, 0), (-1, -1), (2 ** 40, -2 ** 40), (-2 ** 62, 2 ** 62), (123456789, -987654321)]
    cases += [(rng.randrange(-2 ** 62, 2 ** 62), rng.randrange(-2 ** 62, 2 ** 62)) for _ in range(3000)]
    cases += [(rng.randrange(-5000, 5000), rng.randrange(-5000, 5000)) for _ in range(3000)]
    for T in (1, 2, 3, 512, 2006, 200006, 2002006, 2 ** 31 - 1, 2 ** 32 - 1, 4000000007 % 2 ** 32):
        magic = M64 // T
        for xi, yi in cases:
            h = ((xi * 92837111) & M64) ^ ((yi * 689287499) & M64)
            q = (h * magic) >> 64
            r = h - q * T
            assert 0 <= r < 3 * T
            if r >= T:
                r -= T
            if r >= T:
                r -= T
            assert r == h % T, (T, xi, yi)
