"""Hash definitions of SURVEY.md Appendix F.1 (used to compare against the golden vectors taken from the
reference's prebuilt wasm)."""
import numpy as np

FNV_BASIS = 0x14650FB0739D0383  # note: not the standard FNV basis
FNV_PRIME = 0x100000001B3
MASK = (1 << 64) - 1


def fnv_words(words: np.ndarray) -> int:
    """64-bit FNV-1a over u32 words, each fed as 4 bytes least-significant first."""
    h = FNV_BASIS
    for b in np.ascontiguousarray(words, dtype="<u4").tobytes():
        h = ((h ^ b) * FNV_PRIME) & MASK
    return h


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def state_hash(bodies: dict) -> int:
    n = len(bodies["id"])
    w = np.empty((n, 7), np.uint32)
    w[:, 0] = bodies["id"]
    w[:, 1:3] = _bits(bodies["pos"]).reshape(n, 2)
    w[:, 3] = _bits(bodies["angle"])
    w[:, 4:6] = _bits(bodies["momentum"]).reshape(n, 2)
    w[:, 6] = _bits(bodies["ang_momentum"])
    return fnv_words(w.ravel())


def aabb_hash(bodies: dict) -> int:
    n = len(bodies["id"])
    w = np.empty((n, 5), np.uint32)
    w[:, 0] = bodies["id"]
    w[:, 1:5] = _bits(bodies["aabb"]).reshape(n, 4)
    return fnv_words(w.ravel())


def pairs_hash(pairs: np.ndarray) -> int:
    return fnv_words(np.ascontiguousarray(pairs, dtype=np.uint32).ravel())


def manifolds_hash(man: np.ndarray) -> int:
    """[ref_id, inc_id, bits(normal.x), bits(normal.y), n_points, then per point bits(pos.x), bits(pos.y), bits(depth)]"""
    n = len(man)
    if n == 0:
        return FNV_BASIS
    w = np.zeros((n, 11), np.uint32)
    w[:, 0], w[:, 1] = man["ref_id"], man["inc_id"]
    w[:, 2], w[:, 3] = _bits(man["normal_x"]), _bits(man["normal_y"])
    w[:, 4] = man["n_points"]
    for k in range(2):
        w[:, 5 + 3 * k] = _bits(man["pos_x"][:, k])
        w[:, 6 + 3 * k] = _bits(man["pos_y"][:, k])
        w[:, 7 + 3 * k] = _bits(man["depth"][:, k])
    keep = np.zeros((n, 11), bool)
    keep[:, :5] = True
    keep[:, 5:8] = (man["n_points"] >= 1)[:, None]
    keep[:, 8:11] = (man["n_points"] >= 2)[:, None]
    return fnv_words(w[keep])
