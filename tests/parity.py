"""Shared parity assertions: a candidate implementation (CUDA path or CPU emulator) against the oracle."""
import os

import numpy as np

from oracle import ORDER_COLORED, ORDER_REFERENCE, OracleSolver


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bodies_equal(got: dict, want: dict, what=""):
    assert np.array_equal(got["id"], want["id"]), f"{what}: ids/iteration order"
    for k in ("pos", "angle", "momentum", "ang_momentum", "aabb"):
        g, w = bits(got[k]), bits(want[k])
        if not np.array_equal(g, w):
            bad = np.argwhere(g.reshape(len(got["id"]), -1) != w.reshape(len(got["id"]), -1))
            i = bad[0][0]
            err = np.max(np.abs(np.asarray(got[k], np.float64) - np.asarray(want[k], np.float64)))
            raise AssertionError(f"{what}: {k} differs on {len(set(bad[:, 0]))} bodies (first slot {i}, id {got['id'][i]}: "
                                 f"{np.asarray(got[k])[i]} vs {np.asarray(want[k])[i]}), max abs err {err:g}")


def manifold_rows(man: np.ndarray) -> np.ndarray:
    n = len(man)
    cols = [np.minimum(man["ref_id"], man["inc_id"]), np.maximum(man["ref_id"], man["inc_id"]), man["ref_id"],
            man["inc_id"], man["normal_id"], man["n_points"], bits(man["normal_x"]), bits(man["normal_y"])]
    for name in ("pos_x", "pos_y", "depth", "ref_rx", "ref_ry", "inc_rx", "inc_ry"):
        for k in range(2):
            v = bits(man[name][:, k]).copy()
            v[man["n_points"] <= k] = 0
            cols.append(v)
    cols.append(man["color"])
    w = np.stack([np.asarray(c, np.uint32) for c in cols], axis=1) if n else np.zeros((0, 23), np.uint32)
    order = np.lexsort((w[:, 1], w[:, 0]))
    return w[order]


def assert_manifolds_equal(got: np.ndarray, want: np.ndarray, what="", colors=True):
    g, w = manifold_rows(got), manifold_rows(want)
    assert len(g) == len(w), f"{what}: manifold count {len(g)} vs {len(w)}"
    if not colors:
        g, w = g[:, :-1], w[:, :-1]
    if not np.array_equal(g, w):
        bad = np.argwhere(g != w)
        r, c = bad[0]
        raise AssertionError(f"{what}: manifolds differ in {len(set(bad[:, 0]))} rows; first pair ({g[r,0]},{g[r,1]}) column {c}: "
                             f"{g[r, c]:#x} vs {w[r, c]:#x}")


def run_parity(make_candidate, build_scene, steps, check_every=1, pre_step=None, what="", sub_steps=None, iters=None,
               dt=None, options=()):
    """Steps candidate and oracle (coloured Gauss-Seidel order) side by side; everything must be bit-identical.
    `options`: (R2D_OPT_*, value) pairs set on both."""
    from resolve2d_b200 import scenes
    cand = make_candidate()
    orc = OracleSolver(2.0, 4, order=ORDER_COLORED)
    for opt, val in options:
        cand.set_option(opt, val)
        orc.set_option(opt, val)
    cfg = build_scene(cand) or {}
    build_scene(orc)
    S = sub_steps if sub_steps is not None else cfg.get("sub_steps", 4)
    I = iters if iters is not None else cfg.get("iters", 4)
    dt = scenes.DT if dt is None else dt
    assert_bodies_equal(cand.read_bodies(), orc.read_bodies(), f"{what} step 0")
    for step in range(1, steps + 1):
        if pre_step:
            pre_step(cand)
            pre_step(orc)
        cand.process(dt, S, I)
        orc.process(dt, S, I)
        if step % check_every == 0 or step == steps:
            tag = f"{what} step {step}"
            gp, wp = cand.read_pairs(), orc.read_pairs()
            assert np.array_equal(gp, wp), f"{tag}: candidate pair set differs ({len(gp)} vs {len(wp)})"
            assert_manifolds_equal(cand.read_manifolds(), orc.read_manifolds(), tag)
            gs, ws = cand.stats(), orc.stats()
            assert (gs.n_pairs, gs.n_manifolds, gs.n_points, gs.n_colors) == \
                   (ws.n_pairs, ws.n_manifolds, ws.n_points, ws.n_colors), f"{tag}: stats"
            # E (grid entries) is the reference's sum of covered cells only when every body goes through the hashed
            # buckets; the fine grid enters small bodies once (in their home cell) and never more often than that
            if os.environ.get("R2D_BROADPHASE") == "buckets" or os.environ.get("R2D_EMU_BROADPHASE") == "buckets":
                assert gs.n_entries == ws.n_entries, f"{tag}: grid entries"
            else:
                assert gs.n_entries <= ws.n_entries, f"{tag}: grid entries"
            gj, wj = cand.read_joint_order(), orc.read_joint_order()
            assert np.array_equal(gj[0], wj[0]) and np.array_equal(gj[1], wj[1]), f"{tag}: joint order"
            assert_bodies_equal(cand.read_bodies(), orc.read_bodies(), tag)
    return cand, orc
