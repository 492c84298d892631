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
