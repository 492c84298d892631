"""CPU tier: the C-ABI library loads and exports every symbol include/r2d_abi.h declares; host-side registry semantics
(ids, swapRemove, clear) are exercised through the emulator build of the same host code (no GPU compute here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from resolve2d_b200 import _abi
from resolve2d_b200.solver import BodyOptions, DiscOptions, R2DError, RectangleOptions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_abi.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "r2d_abi.h")).read()
    declared = sorted(set(re.findall(r"\b(r2d_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 50
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    assert set(declared) == set(_abi.SIGNATURES), set(declared) ^ set(_abi.SIGNATURES)
    _abi.bind(lib)
    assert lib.r2d_abi_version() == 2


def test_no_device_is_an_error_not_a_fallback():
    """Without a GPU the product must fail loudly (this test only runs where no device exists)."""
    lib = _abi.load_library()
    n = ctypes.c_int(-1)
    assert lib.r2d_device_count(ctypes.byref(n)) == 0
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    assert lib.r2d_create(2.0, 4, 0, ctypes.byref(h)) == -5          # R2D_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.r2d_last_error()


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_abi.BodyOpts) == 36 and ctypes.sizeof(_abi.BodyDesc) == 52
    assert ctypes.sizeof(_abi.JointParams) == 12 and ctypes.sizeof(_abi.BodyState) == 84
    assert ctypes.sizeof(_abi.Manifold) == 84 and ctypes.sizeof(_abi.StepStats) == 48
    assert _abi.body_desc_dtype().itemsize == 52 and _abi.manifold_dtype().itemsize == 84


def test_registry_semantics_ids_swapremove_clear():
    from emu import EmuSolver
    s = EmuSolver(2.0, 4)
    fac = s.entity_factory()
    hs = [fac.make_disc_body(BodyOptions(pos=(i, 0), mass=1), DiscOptions(0.4)) for i in range(5)]
    assert [h.id for h in hs] == [0, 1, 2, 3, 4]                       # lib.zig:66-71
    s.remove_rigid_body(1)                                             # swapRemove: last moves into the hole
    assert [s.body_id_at(i) for i in range(4)] == [0, 4, 2, 3]
    with pytest.raises(R2DError) as e:
        s.remove_rigid_body(1)
    assert e.value.name == "NoSuchIdExists"
    with pytest.raises(R2DError):
        s.body_handle(1).get()
    s.clear()                                                          # lib.zig:181-187: id counter survives (Q16)
    assert s.num_bodies() == 0
    assert fac.make_rectangle_body(BodyOptions(pos=(0, 0), density=1), RectangleOptions()).id == 5
    st = s.body_handle(5).get()
    assert (st.shape_a, st.shape_b, st.mu) == (1.0, 0.5, 0.5)          # RectangleOptions / BodyOptions defaults
    assert st.mass == 0.5 and abs(st.inertia - 0.5 * (1 + 0.25) / 12) < 1e-7


def test_body_options_require_exactly_one_mass_prop():
    with pytest.raises(ValueError):
        BodyOptions(pos=(0, 0)).to_c()
    with pytest.raises(ValueError):
        BodyOptions(pos=(0, 0), mass=1, density=1).to_c()
