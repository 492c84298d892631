"""The reference's flat wasm C ABI (src/wasm_root.zig:18-251, 43 exports over one global solver) re-exported by
resolve2d_b200/csrc/r2d_wasm_compat.c: driven exactly as demos/web/src/wasm_bridge.ts:43-81 drives the wasm module —
solverInit, setup_0_*, then per frame solverProcess and, per body, getRigidBodyPtrFromId + the getters — and compared with
the golden vectors taken from the reference's own binary (tests/golden/wasm_golden.json).

Step 0 (scene construction) is compared for both example scenes, every later step on the impact-free prefix of 0_3, where
the Gauss-Seidel order cannot matter yet (colour order vs the reference's list order: DESIGN.md section 2).
CPU tier: the shim bound to the emulator backend.  GPU tier: the shipped libr2d_wasm_compat.so over libr2d_b200.so."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

import hashing as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wasm_golden.json")))

F32_GETTERS = ["getRigidBodyAABBPosX", "getRigidBodyAABBPosY", "getRigidBodyAABBHalfWidth", "getRigidBodyAABBHalfHeight",
               "getRigidBodyPosX", "getRigidBodyPosY", "getRigidBodyMomentumX", "getRigidBodyMomentumY",
               "getRigidBodyAngularVelocity", "getRigidBodyForceX", "getRigidBodyForceY", "getRigidBodyMass",
               "getRigidBodyAngle", "getRigidBodyAngularMomentum", "getRigidBodyTorque", "getRigidBodyInertia",
               "getRigidBodyFrictionCoeff", "getDiscBodyRadiusAssumeType", "getRectangleBodyWidthAssumeType",
               "getRectangleBodyHeightAssumeType"]
F32_SETTERS = ["setRigidBodyMomentumX", "setRigidBodyMomentumY", "setRigidBodyForceX", "setRigidBodyForceY",
               "setRigidBodyAngularMomentum", "setRigidBodyTorque"]


def reference_export_names():
    """The 43 names, from the committed list (the reference tree is not available on the GPU box)."""
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wasm_exports.json")))


def bind(path):
    lib = C.CDLL(path)
    lib.solverInit.restype, lib.solverInit.argtypes = C.c_bool, [C.c_float, C.c_size_t]
    lib.solverDeinit.restype = None
    lib.solverProcess.restype, lib.solverProcess.argtypes = C.c_bool, [C.c_float, C.c_size_t, C.c_size_t]
    lib.solverGetNumBodies.restype = C.c_size_t
    lib.solverGetBodyIdBasedOnIter.restype, lib.solverGetBodyIdBasedOnIter.argtypes = C.c_uint16, [C.c_size_t]
    lib.solverRemoveBodyById.restype, lib.solverRemoveBodyById.argtypes = C.c_bool, [C.c_uint16]
    lib.solverGetRigidbodyPtrById.restype, lib.solverGetRigidbodyPtrById.argtypes = C.c_size_t, [C.c_uint16]
    lib.getRigidBodyPtrFromId.restype, lib.getRigidBodyPtrFromId.argtypes = C.c_size_t, [C.c_uint16]
    lib.getRigidBodyIdFromPtr.restype, lib.getRigidBodyIdFromPtr.argtypes = C.c_uint16, [C.c_size_t]
    lib.isRigidBodyStatic.restype, lib.isRigidBodyStatic.argtypes = C.c_bool, [C.c_size_t]
    lib.getRigidBodyNumNormals.restype, lib.getRigidBodyNumNormals.argtypes = C.c_size_t, [C.c_size_t]
    lib.getRigidBodyType.restype, lib.getRigidBodyType.argtypes = C.c_size_t, [C.c_size_t]
    lib.getRigidBodyImplementation.restype, lib.getRigidBodyImplementation.argtypes = C.c_uint32, [C.c_size_t]
    for n in ("setup_0_1_car_platformer", "setup_0_2_bridge_stress", "setup_0_3_many_boxes", "setup_0_4_also_many_boxes"):
        getattr(lib, n).restype = C.c_bool
    for n in F32_GETTERS:
        getattr(lib, n).restype, getattr(lib, n).argtypes = C.c_float, [C.c_size_t]
    for n in F32_SETTERS:
        getattr(lib, n).restype, getattr(lib, n).argtypes = None, [C.c_size_t, C.c_float]
    return lib


def frame(lib):
    """One frame of wasm_bridge.ts: every body through getRigidBodyPtrFromId and the getters."""
    n = lib.solverGetNumBodies()
    b = {"id": np.empty(n, np.uint32), "pos": np.empty((n, 2), np.float32), "angle": np.empty(n, np.float32),
         "momentum": np.empty((n, 2), np.float32), "ang_momentum": np.empty(n, np.float32), "aabb": np.empty((n, 4), np.float32)}
    for i in range(n):
        bid = lib.solverGetBodyIdBasedOnIter(i)
        ptr = lib.getRigidBodyPtrFromId(bid)
        assert lib.getRigidBodyIdFromPtr(ptr) == bid and lib.solverGetRigidbodyPtrById(bid) == ptr
        b["id"][i] = bid
        b["pos"][i] = lib.getRigidBodyPosX(ptr), lib.getRigidBodyPosY(ptr)
        b["angle"][i] = lib.getRigidBodyAngle(ptr)
        b["momentum"][i] = lib.getRigidBodyMomentumX(ptr), lib.getRigidBodyMomentumY(ptr)
        b["ang_momentum"][i] = lib.getRigidBodyAngularMomentum(ptr)
        b["aabb"][i] = (lib.getRigidBodyAABBPosX(ptr), lib.getRigidBodyAABBPosY(ptr), lib.getRigidBodyAABBHalfWidth(ptr),
                        lib.getRigidBodyAABBHalfHeight(ptr))
    return b


def drive(lib, last_step):
    assert lib.solverInit(2.0, 4) and lib.solverInit(2.0, 4)       # the second call is a no-op (wasm_root.zig:19)
    try:
        # ---- scene 0_1: construction only (its joints couple bodies from step 1 on, where colour order != list order)
        assert lib.setup_0_2_bridge_stress() and lib.setup_0_4_also_many_boxes()   # empty in the reference
        assert lib.solverGetNumBodies() == 0
        assert lib.setup_0_1_car_platformer()
        g = {r["step"]: r for r in GOLD["0_1_car_platformer"]["steps"]}
        b = frame(lib)
        assert len(b["id"]) == g[0]["n"] == 111
        assert f"{H.state_hash(b):016x}" == g[0]["state"] and f"{H.aabb_hash(b):016x}" == g[0]["aabb"]
        # shape / material getters on the car body (utils.zig:10-13: rectangle 5 x 0.9, density 1, mu 0.4) and a wheel
        p = lib.getRigidBodyPtrFromId(3)
        assert lib.getRigidBodyType(p) == 1 and lib.getRigidBodyNumNormals(p) == 4 and not lib.isRigidBodyStatic(p)
        impl = lib.getRigidBodyImplementation(p)
        assert lib.getRectangleBodyWidthAssumeType(impl) == 5.0 and lib.getRectangleBodyHeightAssumeType(impl) == np.float32(0.9)
        assert lib.getRigidBodyMass(p) == np.float32(5.0) * np.float32(0.9) and lib.getRigidBodyFrictionCoeff(p) == np.float32(0.4)
        w = lib.getRigidBodyPtrFromId(4)
        assert lib.getRigidBodyType(w) == 0 and lib.getRigidBodyNumNormals(w) == 1
        assert lib.getDiscBodyRadiusAssumeType(lib.getRigidBodyImplementation(w)) == 1.0
        assert lib.isRigidBodyStatic(lib.getRigidBodyPtrFromId(0))
        # setters as the key handlers of the demos use them (demos/native/src/main.zig:124-133)
        lib.setRigidBodyAngularMomentum(w, -20.0)
        lib.setRigidBodyTorque(p, 400.0)
        lib.setRigidBodyForceX(p, 3.0)
        lib.setRigidBodyMomentumY(p, 1.5)
        assert lib.getRigidBodyAngularMomentum(w) == -20.0 and lib.getRigidBodyTorque(p) == 400.0
        assert lib.getRigidBodyForceX(p) == 3.0 and lib.getRigidBodyForceY(p) == 0.0 and lib.getRigidBodyMomentumY(p) == 1.5
        assert lib.getRigidBodyAngularVelocity(w) == np.float32(-20.0) / np.float32(lib.getRigidBodyInertia(w))
        assert lib.solverProcess(np.float32(1) / np.float32(60), 4, 4)
        assert lib.getRigidBodyTorque(p) == 0.0 and lib.getRigidBodyForceX(p) == 0.0     # consumed by the step (Q9)
        assert lib.solverRemoveBodyById(110) and not lib.solverRemoveBodyById(110)
        assert lib.solverGetNumBodies() == 110
        lib.solverDeinit()
        assert lib.solverGetNumBodies() == 0 and not lib.solverProcess(0.01, 1, 1)
        # ---- scene 0_3: every frame of the impact-free prefix against the reference binary
        assert lib.solverInit(2.0, 4) and lib.setup_0_3_many_boxes()
        g = {r["step"]: r for r in GOLD["0_3_many_boxes"]["steps"]}
        for step in range(0, last_step + 1):
            if step > 0:
                assert lib.solverProcess(np.float32(1) / np.float32(60), 4, 4)
            b = frame(lib)
            assert len(b["id"]) == g[step]["n"] == 523
            assert f"{H.state_hash(b):016x}" == g[step]["state"], f"0_3 step {step}: state hash"
            assert f"{H.aabb_hash(b):016x}" == g[step]["aabb"], f"0_3 step {step}: aabb hash"
    finally:
        lib.solverDeinit()


def test_shim_exports_every_name_of_the_reference():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "all"], stdout=subprocess.DEVNULL)
    names = reference_export_names()
    assert len(names) == 43
    src = open(os.path.join(ROOT, "resolve2d_b200", "csrc", "r2d_wasm_compat.c")).read()
    defined = set(re.findall(r"^EXPORT [^\n(]*?\b([A-Za-z_0-9]+)\(", src, re.M))
    assert set(names) <= defined, sorted(set(names) - defined)
    for path in (os.path.join(ROOT, "tests", "emu", "_build", "libr2d_wasm_compat_emu.so"),
                 os.path.join(ROOT, "resolve2d_b200", "libr2d_wasm_compat.so")):
        if os.path.exists(path):
            syms = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
            missing = [n for n in names if not re.search(rf"\bT {n}$", syms, re.M)]
            assert not missing, (path, missing)


def test_shim_over_the_emulator_matches_the_reference_binary():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "all"], stdout=subprocess.DEVNULL)
    drive(bind(os.path.join(ROOT, "tests", "emu", "_build", "libr2d_wasm_compat_emu.so")), last_step=40)


@pytest.mark.gpu
def test_gpu_shim_matches_the_reference_binary():
    drive(bind(os.path.join(ROOT, "resolve2d_b200", "libr2d_wasm_compat.so")), last_step=40)
