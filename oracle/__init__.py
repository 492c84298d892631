"""TEST INFRASTRUCTURE — CPU oracle of resolve2d's Solver.process (see oracle/r2d_oracle.cpp).

`OracleSolver` exposes the same Python surface as `resolve2d_b200.Solver`, backed by the single-threaded C++
restatement in oracle/_build/libr2d_oracle.so.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this package; the product (resolve2d_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from resolve2d_b200 import _abi
from resolve2d_b200.solver import Solver

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libr2d_oracle.so")
ORDER_REFERENCE, ORDER_COLORED = 0, 1
_lib = None

_ORACLE_SYMBOLS = [
    "create", "destroy", "clear", "set_mode", "make_disc", "make_rect", "make_bodies", "make_gravity",
    "make_distance_joint", "make_offset_distance_joint", "make_fixed_position_joint", "make_motor_joint",
    "exclude_pair", "remove_body", "process", "step", "synchronize", "reorder", "set_reorder_interval", "num_bodies", "body_id_at", "body_get",
    "body_set_static", "body_set_pos", "body_set_angle", "body_set_momentum", "body_set_ang_momentum",
    "body_set_force", "body_set_torque", "read_bodies", "write_forces", "read_pairs", "read_manifolds",
    "read_joint_order", "get_stats",
]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile recipe (g++, -ffp-contract=off)."""
    src = os.path.join(_HERE, "r2d_oracle.cpp")
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "_build/libr2d_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        sigs = {"r2d_" + n: _abi.SIGNATURES["r2d_" + n] for n in _ORACLE_SYMBOLS}
        _abi.bind(lib, sigs, "r2d_", "orc_")
        lib.orc_set_gs_order.restype = C.c_int
        lib.orc_set_gs_order.argtypes = [C.c_void_p, C.c_int]
        lib.orc_set_option.restype = C.c_int
        lib.orc_set_option.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        lib.orc_timed_steps.restype = C.c_double
        lib.orc_timed_steps.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32]
        lib.orc_load_state.restype = C.c_int
        lib.orc_load_state.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 5
        lib.orc_raw_candidates.restype = C.c_uint64
        lib.orc_raw_candidates.argtypes = [C.c_void_p]
        lib.orc_sinf.restype = C.c_float
        lib.orc_sinf.argtypes = [C.c_float]
        lib.orc_cosf.restype = C.c_float
        lib.orc_cosf.argtypes = [C.c_float]
        lib.orc_aabb_intersects.restype = C.c_int
        lib.orc_aabb_intersects.argtypes = [C.c_float] * 8
        lib.orc_cell_hash.restype = C.c_uint64
        lib.orc_cell_hash.argtypes = [C.c_uint64, C.c_int64, C.c_int64]
        _lib = lib
    return _lib


class OracleSolver(Solver):
    """CPU oracle with the Solver/EntityFactory surface.  `order`: ORDER_REFERENCE = the reference's insertion-order
    Gauss-Seidel sweep; ORDER_COLORED = the colour order the CUDA path sweeps in (same arithmetic, other order)."""

    _prefix = "orc_"

    def __init__(self, spatialhash_cell_width: float = 2.0, spatialhash_table_size_mult: int = 4,
                 order: int = ORDER_REFERENCE):
        super().__init__(spatialhash_cell_width, spatialhash_table_size_mult, 0, _lib=load())
        self.set_gs_order(order)

    def set_gs_order(self, order: int):
        self._lib.orc_set_gs_order(self._h, order)

    def set_option(self, option: int, value: int):
        assert self._lib.orc_set_option(self._h, option, value) == 0

    def set_stream(self, cuda_stream):  # no device
        raise NotImplementedError

    def load_state(self, bodies: dict):
        """Overwrite pos/angle/momentum/ang_momentum/aabb of all bodies (iteration order) from a read_bodies() dict."""
        import numpy as np
        a = {k: np.ascontiguousarray(bodies[k], dtype=np.float32) for k in ("pos", "angle", "momentum", "ang_momentum", "aabb")}
        st = self._lib.orc_load_state(self._h, len(a["angle"]), a["pos"].ctypes.data, a["angle"].ctypes.data,
                                      a["momentum"].ctypes.data, a["ang_momentum"].ctypes.data, a["aabb"].ctypes.data)
        assert st == 0, st

    def timed_steps(self, dt: float, sub_steps: int, iters: int, steps: int) -> float:
        return self._lib.orc_timed_steps(self._h, dt, sub_steps, iters, steps)

    def raw_candidates(self) -> int:
        return self._lib.orc_raw_candidates(self._h)
