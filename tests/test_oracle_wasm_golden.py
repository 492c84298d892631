"""Pins the CPU oracle against outputs of THE REFERENCE ITSELF: tests/golden/wasm_golden.json was produced by executing
the reference's shipped binary (demos/web/public/resolve2d.wasm) in oracle/wasm_interp.cpp (generator:
tests/golden/make_wasm_golden.py).  Per run: state and AABB hashes after EVERY process() call plus full raw body dumps
at a few steps, for both example scenes, the key-driven variant, other (dt, sub_steps, iters) settings and body removal
(swapRemove iteration order).  Everything is compared bit for bit."""
import json
import os

import numpy as np
import pytest

import hashing as H
from oracle import ORDER_REFERENCE, OracleSolver
from resolve2d_b200 import scenes

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wasm_golden.json")))
RUNS = [k for k in GOLD if not k.startswith("_")]
f32 = np.float32


def _hex(v):
    return f"{v:016x}"


@pytest.mark.parametrize("name", RUNS)
def test_oracle_matches_the_executed_reference_binary(name):
    g = GOLD[name]
    s = OracleSolver(2.0, 4, order=ORDER_REFERENCE)
    (scenes.setup_0_1_car_platformer if g["scene"] == "0_1" else scenes.setup_0_3_many_boxes)(s)
    dt = f32(1.0) / f32(g["dt_rate"])
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
        assert _hex(H.state_hash(b)) == rec["state"], f"{name} step {step}: state hash"
        assert _hex(H.aabb_hash(b)) == rec["aabb"], f"{name} step {step}: aabb hash"
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


def test_wasm_golden_agrees_with_survey_appendix_f():
    """The vectors regenerated here from the binary equal the ones the survey recorded (Appendix F)."""
    af = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_f.json")))
    for name in ("0_3_many_boxes", "0_1_car_platformer", "0_1_car_platformer_driven"):
        by_step = {r["step"]: r for r in GOLD[name]["steps"]}
        assert by_step[0]["state"] == af[name]["step0"]["state"] and by_step[0]["aabb"] == af[name]["step0"]["aabb"]
        for r in af[name]["steps"]:
            assert by_step[r["step"]]["state"] == r["state"], (name, r["step"])
            assert by_step[r["step"]]["aabb"] == r["aabb"], (name, r["step"])
