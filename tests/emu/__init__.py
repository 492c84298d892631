"""TEST TOOLING — `EmuSolver`/`EmuBatch`: the Solver surface backed by tests/emu/_build/libr2d_emu.so, a serial CPU
driver of the CUDA kernels' per-thread bodies (see emu.cpp).  Used only by the CPU-only test tier."""
import ctypes as C
import os
import subprocess

from resolve2d_b200 import _abi
from resolve2d_b200.solver import Batch, ShardedBatch, Solver

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libr2d_emu.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        lib = C.CDLL(LIB_PATH)
        _abi.bind(lib, _abi.SIGNATURES, "r2d_", "emu_")
        _lib = lib
    return _lib


class EmuSolver(Solver):
    _prefix = "emu_"

    def __init__(self, cell_width=2.0, table_mult=4, **kw):
        if "_lib" not in kw:
            kw["_lib"] = load()
        super().__init__(cell_width, table_mult, 0, **kw)


class EmuBatch:
    def __init__(self, n_worlds, cell_width=2.0, table_mult=4):
        self._lib = load()
        h = C.c_void_p()
        assert self._lib.emu_batch_create(n_worlds, cell_width, table_mult, 0, C.byref(h)) == 0
        self._h = h

    def world(self, w):
        h = C.c_void_p()
        assert self._lib.emu_batch_world(self._h, w, C.byref(h)) == 0
        return EmuSolver(_lib=self._lib, _handle=h, _owner=self)

    def process(self, dt, s, i):
        st = self._lib.emu_batch_process(self._h, dt, s, i)
        if st != 0:
            raise _abi.R2DError(st, "emu_batch_process")

    def reorder(self):
        st = self._lib.emu_batch_reorder(self._h)
        if st != 0:
            raise _abi.R2DError(st, "emu_batch_reorder")

    def set_reorder_interval(self, steps):
        assert self._lib.emu_batch_set_reorder_interval(self._h, C.c_uint32(steps)) == 0

    def __del__(self):
        try:
            self._lib.emu_batch_destroy(self._h)
        except Exception:
            pass


class EmuShardedBatch(ShardedBatch):
    """r2d_sharded_* over the CPU emulator backend: the world -> shard map, the per-shard host threads and the world-major
    slicing of the bulk arrays are the product's own code (r2d_capi.inc); only the kernels are emulated."""
    _prefix = "emu_"
    _world_class = EmuSolver

    def __init__(self, n_worlds, n_shards, cell_width=2.0, table_mult=4):
        super().__init__(n_worlds, [0] * n_shards, cell_width, table_mult, _lib=load())
