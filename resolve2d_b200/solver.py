"""Host-side mirror of the reference's `Solver` + `EntityFactory` (src/core/lib.zig:13-315) over the C ABI.

Same names, argument meaning and error behaviour as the Zig API, so tests read like tests of the reference:

    solver = Solver(2.0, 4)                      # Solver.init(alloc, cell_width, table_size_mult)   lib.zig:145
    fac = solver.entity_factory()                # solver.entityFactory()                            lib.zig:312
    fac.make_downwards_gravity(9.82)             # makeDownwardsGravity                              lib.zig:95
    h = fac.make_rectangle_body(BodyOptions(pos=(0, -5), density=5, mu=0.3), RectangleOptions(1000, 10))
    h.set_static(True)                           # h.body_unwrap().static = true
    solver.process(1 / 60, 4, 4)                 # solver.process(alloc, dt, sub_steps, collision_iters)  lib.zig:189

Raw `*RigidBody` pointers (lib.zig:23-31) do not exist — the body state lives in HBM as structure-of-arrays — so
field access goes through `BodyHandle.get()/set_*()` or the bulk `read_bodies()`.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _abi
from ._abi import (BodyDesc, BodyOpts, BodyState, JointParams, Manifold, R2DError, StepStats, body_desc_dtype,
                   manifold_dtype)


@dataclass
class BodyOptions:
    """EntityFactory.BodyOptions (lib.zig:35-46); give exactly one of density / mass."""
    pos: Tuple[float, float] = (0.0, 0.0)
    vel: Tuple[float, float] = (0.0, 0.0)
    angle: float = 0.0
    omega: float = 0.0
    mu: float = 0.5
    density: Optional[float] = None
    mass: Optional[float] = None

    def to_c(self) -> BodyOpts:
        if (self.density is None) == (self.mass is None):
            raise ValueError("BodyOptions.mass_prop: give exactly one of density= or mass=")
        is_density = self.density is not None
        return BodyOpts(self.pos[0], self.pos[1], self.vel[0], self.vel[1], self.angle, self.omega, self.mu,
                        self.density if is_density else self.mass, 1 if is_density else 0)


@dataclass
class DiscOptions:
    radius: float = 1.0       # lib.zig:48-50


@dataclass
class RectangleOptions:
    width: float = 1.0        # lib.zig:52-55
    height: float = 0.5


@dataclass
class Parameters:
    """Constraint.Parameters (Constraints/Constraint.zig:27-31)."""
    power_max: float = math.inf
    power_min: float = -math.inf
    beta: float = 10.0

    def to_c(self) -> JointParams:
        return JointParams(self.power_max, self.power_min, self.beta)


class BodyHandle:
    """EntityFactory.BodyHandle (lib.zig:15-32): an id plus its solver."""

    def __init__(self, solver: "Solver", id: int):
        self.solver = solver
        self.id = id

    def get(self) -> BodyState:
        """`body_unwrap().*` by value; raises NoSuchIdExists where the reference would panic."""
        st = BodyState()
        self.solver._check(self.solver._fn("body_get")(self.solver._h, self.id, C.byref(st)), "body_get")
        return st

    body = get
    body_unwrap = get

    def set_static(self, v: bool = True):
        self.solver._check(self.solver._fn("body_set_static")(self.solver._h, self.id, 1 if v else 0), "set_static")

    def set_pos(self, x: float, y: float):
        self.solver._check(self.solver._fn("body_set_pos")(self.solver._h, self.id, x, y), "set_pos")

    def set_angle(self, a: float):
        self.solver._check(self.solver._fn("body_set_angle")(self.solver._h, self.id, a), "set_angle")

    def set_momentum(self, x: float, y: float):
        self.solver._check(self.solver._fn("body_set_momentum")(self.solver._h, self.id, x, y), "set_momentum")

    def set_ang_momentum(self, l: float):
        self.solver._check(self.solver._fn("body_set_ang_momentum")(self.solver._h, self.id, l), "set_ang_momentum")

    def set_force(self, x: float, y: float):
        self.solver._check(self.solver._fn("body_set_force")(self.solver._h, self.id, x, y), "set_force")

    def set_torque(self, t: float):
        self.solver._check(self.solver._fn("body_set_torque")(self.solver._h, self.id, t), "set_torque")


class EntityFactory:
    """EntityFactory (lib.zig:13-130)."""

    def __init__(self, solver: "Solver"):
        self.solver = solver

    def make_disc_body(self, bo: BodyOptions, go: DiscOptions = DiscOptions()) -> BodyHandle:  # lib.zig:73
        s = self.solver
        out = C.c_uint32()
        o = bo.to_c()
        s._check(s._fn("make_disc")(s._h, C.byref(o), go.radius, C.byref(out)), "makeDiscBody")
        return BodyHandle(s, out.value)

    def make_rectangle_body(self, bo: BodyOptions, go: RectangleOptions = RectangleOptions()) -> BodyHandle:  # :84
        s = self.solver
        out = C.c_uint32()
        o = bo.to_c()
        s._check(s._fn("make_rect")(s._h, C.byref(o), go.width, go.height, C.byref(out)), "makeRectangleBody")
        return BodyHandle(s, out.value)

    def make_bodies(self, descs: np.ndarray) -> int:
        """Bulk creation from a structured array of dtype `body_desc_dtype()`; returns the first id."""
        s = self.solver
        descs = np.ascontiguousarray(descs, dtype=body_desc_dtype())
        assert descs.dtype.itemsize == C.sizeof(BodyDesc)
        out = C.c_uint32()
        s._check(s._fn("make_bodies")(s._h, descs.ctypes.data, descs.shape[0], C.byref(out)), "make_bodies")
        return out.value

    def make_downwards_gravity(self, g: float):  # lib.zig:95
        s = self.solver
        s._check(s._fn("make_gravity")(s._h, g), "makeDownwardsGravity")

    def make_offset_distance_joint(self, params: Parameters, h1: BodyHandle, h2: BodyHandle, r1, r2,
                                   target_distance: float) -> int:  # lib.zig:100
        s = self.solver
        out = C.c_size_t()
        p = params.to_c()
        s._check(s._fn("make_offset_distance_joint")(s._h, C.byref(p), h1.id, h2.id, r1[0], r1[1], r2[0], r2[1],
                                                      target_distance, C.byref(out)), "makeOffsetDistanceJoint")
        return out.value

    def make_distance_joint(self, params: Parameters, h1: BodyHandle, h2: BodyHandle, target_distance: float) -> int:
        s = self.solver
        out = C.c_size_t()
        p = params.to_c()
        s._check(s._fn("make_distance_joint")(s._h, C.byref(p), h1.id, h2.id, target_distance, C.byref(out)),
                 "makeDistanceJoint")
        return out.value

    def make_fixed_position_joint(self, params: Parameters, h: BodyHandle, target_position) -> int:  # lib.zig:112
        s = self.solver
        out = C.c_size_t()
        p = params.to_c()
        s._check(s._fn("make_fixed_position_joint")(s._h, C.byref(p), h.id, target_position[0], target_position[1],
                                                     C.byref(out)), "makeFixedPositionJoint")
        return out.value

    def make_motor_joint(self, params: Parameters, h: BodyHandle, target_omega: float) -> int:  # lib.zig:118
        s = self.solver
        out = C.c_size_t()
        p = params.to_c()
        s._check(s._fn("make_motor_joint")(s._h, C.byref(p), h.id, target_omega, C.byref(out)), "makeMotorJoint")
        return out.value

    def exclude_collision_pair(self, h1: BodyHandle, h2: BodyHandle):  # lib.zig:124
        s = self.solver
        s._check(s._fn("exclude_pair")(s._h, h1.id, h2.id), "excludeCollisionPair")


class Solver:
    """Solver (lib.zig:132-315) on one B200.  `cell_width` / `table_mult` are stored but, exactly like the reference
    (lib.zig:254-255), ignored unless `set_mode(MODE_FAST)` is selected."""

    _prefix = "r2d_"

    def __init__(self, spatialhash_cell_width: float = 2.0, spatialhash_table_size_mult: int = 4, device: int = 0, *,
                 _lib=None, _handle=None, _owner=None):
        self._lib = _lib if _lib is not None else _abi.load_library()
        self._owner = _owner  # a Batch keeps its worlds alive; borrowed handles are not destroyed here
        if _handle is not None:
            self._h = _handle
        else:
            h = C.c_void_p()
            self._check(self._fn("create")(spatialhash_cell_width, spatialhash_table_size_mult, device, C.byref(h)),
                        "Solver.init")
            self._h = h

    # -- plumbing -----------------------------------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)

    def _check(self, status, what):
        if status != 0:
            detail = ""
            if self._prefix == "r2d_":
                detail = (self._lib.r2d_last_error() or b"").decode("utf-8", "replace")
            raise R2DError(status, what, detail)

    def deinit(self):  # lib.zig:159
        if getattr(self, "_h", None) is not None and self._owner is None:
            self._fn("destroy")(self._h)
        self._h = None

    def __del__(self):
        try:
            self.deinit()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.deinit()

    # -- reference API -------------------------------------------------------------------------------------------
    def clear(self):  # lib.zig:181
        self._check(self._fn("clear")(self._h), "clear")

    def set_mode(self, mode: int):
        self._check(self._fn("set_mode")(self._h, mode), "set_mode")

    def set_option(self, option: int, value: int):
        """R2D_OPT_* (warm starting, sleeping: the reference's roadmap items, off by default)."""
        self._check(self._fn("set_option")(self._h, option, value), "set_option")

    def set_stream(self, cuda_stream: int):
        self._check(self._fn("set_stream")(self._h, C.c_void_p(cuda_stream)), "set_stream")

    def reorder(self):
        """Re-derive the spatial device order from the current positions (results are unaffected)."""
        self._check(self._fn("reorder")(self._h), "reorder")

    def set_reorder_interval(self, steps: int):
        self._check(self._fn("set_reorder_interval")(self._h, steps), "set_reorder_interval")

    def process(self, dt: float, sub_steps: int, collision_iters: int):  # lib.zig:189
        self._check(self._fn("process")(self._h, dt, sub_steps, collision_iters), "process")

    step = process  # north_star's name for the same call

    def process_read(self, dt: float, sub_steps: int, collision_iters: int, out: dict) -> dict:
        """process() + read_bodies(out) in one call and one host synchronisation (r2d_process_read): the export of the
        new state is enqueued behind the step's last kernel.  `out` holds preallocated (ideally pinned) arrays or None
        per key, as for read_bodies."""
        def p(k):
            a = out.get(k)
            return None if a is None else C.c_void_p(a.ctypes.data)
        self._check(self._fn("process_read")(self._h, dt, sub_steps, collision_iters, p("id"), p("pos"), p("angle"),
                                             p("momentum"), p("ang_momentum"), p("aabb"), self.num_bodies()), "process_read")
        return out

    def synchronize(self):
        self._check(self._fn("synchronize")(self._h), "synchronize")

    def body_handle(self, id: int) -> BodyHandle:  # lib.zig:301
        return BodyHandle(self, id)

    def remove_rigid_body(self, id: int):  # lib.zig:308 -> error.NoSuchIdExists
        self._check(self._fn("remove_body")(self._h, id), "removeRigidBody")

    def entity_factory(self) -> EntityFactory:  # lib.zig:312
        return EntityFactory(self)

    # -- state access (wasm_root.zig getters) ------------------------------------------------------------------------
    def num_bodies(self) -> int:
        n = C.c_size_t()
        self._check(self._fn("num_bodies")(self._h, C.byref(n)), "num_bodies")
        return n.value

    def body_id_at(self, i: int) -> int:
        out = C.c_uint32()
        self._check(self._fn("body_id_at")(self._h, i, C.byref(out)), "body_id_at")
        return out.value

    def read_bodies(self, out: Optional[dict] = None) -> dict:
        """Bulk SoA readback in solver iteration order.  `out` may hold preallocated (e.g. pinned) arrays."""
        n = self.num_bodies()
        if out is None:
            out = {
                "id": np.empty(n, np.uint32), "pos": np.empty((n, 2), np.float32), "angle": np.empty(n, np.float32),
                "momentum": np.empty((n, 2), np.float32), "ang_momentum": np.empty(n, np.float32),
                "aabb": np.empty((n, 4), np.float32),
            }

        def p(k):
            a = out.get(k)
            return None if a is None else C.c_void_p(a.ctypes.data)
        self._check(self._fn("read_bodies")(self._h, p("id"), p("pos"), p("angle"), p("momentum"), p("ang_momentum"),
                                            p("aabb"), n), "read_bodies")
        return out

    def write_forces(self, force_xy_torque: np.ndarray):
        a = np.ascontiguousarray(force_xy_torque, dtype=np.float32)
        self._check(self._fn("write_forces")(self._h, C.c_void_p(a.ctypes.data), a.shape[0]), "write_forces")

    def read_pairs(self) -> np.ndarray:
        """Candidate set of the last process(): sorted unique (lo_id, hi_id) rows."""
        n = C.c_size_t()
        self._check(self._fn("read_pairs")(self._h, None, None, 0, C.byref(n)), "read_pairs")
        lo = np.empty(n.value, np.uint32)
        hi = np.empty(n.value, np.uint32)
        if n.value:
            self._check(self._fn("read_pairs")(self._h, C.c_void_p(lo.ctypes.data), C.c_void_p(hi.ctypes.data), n.value,
                                               C.byref(n)), "read_pairs")
        return np.stack([lo, hi], axis=1)

    def read_manifolds(self) -> np.ndarray:
        n = C.c_size_t()
        self._check(self._fn("read_manifolds")(self._h, None, 0, C.byref(n)), "read_manifolds")
        out = np.zeros(n.value, manifold_dtype())
        assert out.dtype.itemsize == C.sizeof(Manifold)
        if n.value:
            self._check(self._fn("read_manifolds")(self._h, C.c_void_p(out.ctypes.data), n.value, C.byref(n)),
                        "read_manifolds")
        return out

    def read_joint_order(self):
        n = C.c_size_t()
        self._check(self._fn("read_joint_order")(self._h, None, None, 0, C.byref(n)), "read_joint_order")
        idx = np.empty(n.value, np.uint32)
        col = np.empty(n.value, np.uint32)
        if n.value:
            self._check(self._fn("read_joint_order")(self._h, C.c_void_p(idx.ctypes.data), C.c_void_p(col.ctypes.data),
                                                     n.value, C.byref(n)), "read_joint_order")
        return idx, col

    def stats(self) -> StepStats:
        st = StepStats()
        self._check(self._fn("get_stats")(self._h, C.byref(st)), "get_stats")
        return st

    def profile_enable(self, on: bool = True):
        self._check(self._fn("profile_enable")(self._h, 1 if on else 0), "profile_enable")

    def profile_read(self, reset: bool = True):
        ms = (C.c_double * len(_abi.KCLASS_NAMES))()
        cnt = (C.c_uint64 * len(_abi.KCLASS_NAMES))()
        self._check(self._fn("profile_read")(self._h, ms, cnt, 1 if reset else 0), "profile_read")
        return {k: (ms[i], cnt[i]) for i, k in enumerate(_abi.KCLASS_NAMES)}


class Batch:
    """Thousands of independent worlds stepped together on one GPU (north_star "Batched-worlds mode")."""

    def __init__(self, n_worlds: int, spatialhash_cell_width: float = 2.0, spatialhash_table_size_mult: int = 4,
                 device: int = 0):
        self._lib = _abi.load_library()
        h = C.c_void_p()
        _abi.check(self._lib, self._lib.r2d_batch_create(n_worlds, spatialhash_cell_width, spatialhash_table_size_mult,
                                                         device, C.byref(h)), "batch_create")
        self._h = h
        self.n_worlds = n_worlds

    def world(self, w: int) -> Solver:
        h = C.c_void_p()
        _abi.check(self._lib, self._lib.r2d_batch_world(self._h, w, C.byref(h)), "batch_world")
        return Solver(_lib=self._lib, _handle=h, _owner=self)

    def set_mode(self, mode: int):
        _abi.check(self._lib, self._lib.r2d_batch_set_mode(self._h, mode), "batch_set_mode")

    def set_option(self, option: int, value: int):
        _abi.check(self._lib, self._lib.r2d_batch_set_option(self._h, option, value), "batch_set_option")

    def set_stream(self, cuda_stream: int):
        _abi.check(self._lib, self._lib.r2d_batch_set_stream(self._h, C.c_void_p(cuda_stream)), "batch_set_stream")

    def reorder(self):
        _abi.check(self._lib, self._lib.r2d_batch_reorder(self._h), "batch_reorder")

    def set_reorder_interval(self, steps: int):
        _abi.check(self._lib, self._lib.r2d_batch_set_reorder_interval(self._h, steps), "batch_set_reorder_interval")

    def process(self, dt: float, sub_steps: int, collision_iters: int):
        _abi.check(self._lib, self._lib.r2d_batch_process(self._h, dt, sub_steps, collision_iters), "batch_process")

    def process_read(self, dt: float, sub_steps: int, collision_iters: int, out: dict) -> dict:
        """batch_process() + read_bodies(out) of every world in one call and one host synchronisation."""
        def p(k):
            a = out.get(k)
            return None if a is None else C.c_void_p(a.ctypes.data)
        _abi.check(self._lib, self._lib.r2d_batch_process_read(self._h, dt, sub_steps, collision_iters, p("id"), p("pos"),
                                                                p("angle"), p("momentum"), p("ang_momentum"), p("aabb"),
                                                                self.num_bodies()), "batch_process_read")
        return out

    def synchronize(self):
        _abi.check(self._lib, self._lib.r2d_batch_synchronize(self._h), "batch_synchronize")

    def num_bodies(self) -> int:
        n = C.c_size_t()
        _abi.check(self._lib, self._lib.r2d_batch_num_bodies(self._h, C.byref(n)), "batch_num_bodies")
        return n.value

    def read_bodies(self, out: Optional[dict] = None) -> dict:
        n = self.num_bodies()
        if out is None:
            out = {
                "id": np.empty(n, np.uint32), "pos": np.empty((n, 2), np.float32), "angle": np.empty(n, np.float32),
                "momentum": np.empty((n, 2), np.float32), "ang_momentum": np.empty(n, np.float32),
                "aabb": np.empty((n, 4), np.float32),
            }

        def p(k):
            a = out.get(k)
            return None if a is None else C.c_void_p(a.ctypes.data)
        _abi.check(self._lib, self._lib.r2d_batch_read_bodies(self._h, p("id"), p("pos"), p("angle"), p("momentum"),
                                                              p("ang_momentum"), p("aabb"), n), "batch_read_bodies")
        return out

    def write_forces(self, force_xy_torque: np.ndarray):
        a = np.ascontiguousarray(force_xy_torque, dtype=np.float32)
        _abi.check(self._lib, self._lib.r2d_batch_write_forces(self._h, C.c_void_p(a.ctypes.data), a.shape[0]),
                   "batch_write_forces")

    def stats(self) -> StepStats:
        st = StepStats()
        _abi.check(self._lib, self._lib.r2d_batch_get_stats(self._h, C.byref(st)), "batch_get_stats")
        return st

    def profile_enable(self, on: bool = True):
        _abi.check(self._lib, self._lib.r2d_batch_profile_enable(self._h, 1 if on else 0), "batch_profile_enable")

    def profile_read(self, reset: bool = True):
        ms = (C.c_double * len(_abi.KCLASS_NAMES))()
        cnt = (C.c_uint64 * len(_abi.KCLASS_NAMES))()
        _abi.check(self._lib, self._lib.r2d_batch_profile_read(self._h, ms, cnt, 1 if reset else 0), "profile_read")
        return {k: (ms[i], cnt[i]) for i, k in enumerate(_abi.KCLASS_NAMES)}

    def destroy(self):
        if getattr(self, "_h", None) is not None:
            self._lib.r2d_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class ShardedBatch:
    """BASELINE config 5 as a library object: `n_worlds` independent worlds sharded over the GPUs `devices` of one box
    (world w -> shard floor(w * G / n_worlds)), one host thread + one stream per device inside the library,
    state resident per device, no collective on the data path.  Bulk arrays are world-major over all worlds."""

    _prefix = "r2d_"
    _world_class = Solver

    def __init__(self, n_worlds: int, devices: Sequence[int], spatialhash_cell_width: float = 2.0,
                 spatialhash_table_size_mult: int = 4, *, _lib=None):
        self._lib = _lib if _lib is not None else _abi.load_library()
        h = C.c_void_p()
        dev = (C.c_int * len(devices))(*devices)
        self._check(self._fn("sharded_create")(n_worlds, dev, len(devices), spatialhash_cell_width,
                                               spatialhash_table_size_mult, C.byref(h)), "sharded_create")
        self._h = h
        self.n_worlds = n_worlds
        self.n_shards = len(devices)

    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)

    def _check(self, status, what):
        if status != 0:
            detail = ""
            if self._prefix == "r2d_":
                detail = (self._lib.r2d_last_error() or b"").decode("utf-8", "replace")
            raise R2DError(status, what, detail)

    def world(self, w: int) -> Solver:
        h = C.c_void_p()
        self._check(self._fn("sharded_world")(self._h, w, C.byref(h)), "sharded_world")
        return self._world_class(_lib=self._lib, _handle=h, _owner=self)

    def shard(self, k: int):
        """(first world, number of worlds) of shard k."""
        first, n = C.c_uint32(), C.c_uint32()
        self._check(self._fn("sharded_shard")(self._h, k, None, C.byref(first), C.byref(n)), "sharded_shard")
        return first.value, n.value

    def set_mode(self, mode: int):
        self._check(self._fn("sharded_set_mode")(self._h, mode), "sharded_set_mode")

    def reorder(self):
        self._check(self._fn("sharded_reorder")(self._h), "sharded_reorder")

    def process(self, dt: float, sub_steps: int, collision_iters: int):
        self._check(self._fn("sharded_process")(self._h, dt, sub_steps, collision_iters), "sharded_process")

    def process_read(self, dt: float, sub_steps: int, collision_iters: int, out: dict) -> dict:
        def p(k):
            a = out.get(k)
            return None if a is None else C.c_void_p(a.ctypes.data)
        self._check(self._fn("sharded_process_read")(self._h, dt, sub_steps, collision_iters, p("id"), p("pos"), p("angle"),
                                                     p("momentum"), p("ang_momentum"), p("aabb"), self.num_bodies()),
                    "sharded_process_read")
        return out

    def synchronize(self):
        self._check(self._fn("sharded_synchronize")(self._h), "sharded_synchronize")

    def num_bodies(self) -> int:
        n = C.c_size_t()
        self._check(self._fn("sharded_num_bodies")(self._h, C.byref(n)), "sharded_num_bodies")
        return n.value

    def read_bodies(self, out: Optional[dict] = None) -> dict:
        n = self.num_bodies()
        if out is None:
            out = {
                "id": np.empty(n, np.uint32), "pos": np.empty((n, 2), np.float32), "angle": np.empty(n, np.float32),
                "momentum": np.empty((n, 2), np.float32), "ang_momentum": np.empty(n, np.float32),
                "aabb": np.empty((n, 4), np.float32),
            }

        def p(k):
            a = out.get(k)
            return None if a is None else C.c_void_p(a.ctypes.data)
        self._check(self._fn("sharded_read_bodies")(self._h, p("id"), p("pos"), p("angle"), p("momentum"),
                                                    p("ang_momentum"), p("aabb"), n), "sharded_read_bodies")
        return out

    def write_forces(self, force_xy_torque: np.ndarray):
        a = np.ascontiguousarray(force_xy_torque, dtype=np.float32)
        self._check(self._fn("sharded_write_forces")(self._h, C.c_void_p(a.ctypes.data), a.shape[0]), "sharded_write_forces")

    def stats(self) -> StepStats:
        st = StepStats()
        self._check(self._fn("sharded_get_stats")(self._h, C.byref(st)), "sharded_get_stats")
        return st

    def destroy(self):
        if getattr(self, "_h", None) is not None:
            self._fn("sharded_destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
