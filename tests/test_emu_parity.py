"""CPU-only tier: the per-thread kernel bodies (r2d_pipeline.cuh) and the host registry, driven serially by
tests/emu, must reproduce the oracle (coloured Gauss-Seidel order) bit for bit: candidate-pair sets, manifolds,
colours, joint order and body state."""
import numpy as np
import pytest

from emu import EmuBatch, EmuSolver
from oracle import ORDER_COLORED, OracleSolver
from parity import assert_bodies_equal, run_parity
from resolve2d_b200 import scenes


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
    # a plank on 48 discs: more manifolds on one body than its list holds -> rounds, same colours as the oracle
    cand, _ = run_parity(lambda: EmuSolver(2.0, 4), scenes.build_hub, 60, check_every=10, what="hub")
    st = cand.stats()
    assert st.n_colors >= 40 and st.n_color_rounds > 0
