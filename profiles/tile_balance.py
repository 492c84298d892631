"""How evenly do k_solve_tiles's tiles (equal body counts in Morton order) share the manifolds of pile100k?"""
import sys
sys.path.insert(0, '.')
import numpy as np
from resolve2d_b200 import Solver, scenes
s = Solver(2.0, 4); scenes.build_pile100k(s)
for _ in range(230): s.process(scenes.DT, 4, 4)
b = s.read_bodies(); m = s.read_manifolds()
ids = b["id"]; pos = b["pos"].astype(np.float32)
n = len(ids)
mn = pos.min(axis=0)
q = np.clip(((pos - mn) * np.float32(0.5)), 0, 65535).astype(np.uint32)
def spread(v):
    v = v & 0xFFFF
    v = (v | (v << 8)) & 0x00FF00FF; v = (v | (v << 4)) & 0x0F0F0F0F; v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555
    return v
key = spread(q[:, 0]) | (spread(q[:, 1]) << 1)
order = np.argsort(key, kind="stable")
rank = np.empty(n, np.int64); rank[order] = np.arange(n)
rank_of_id = np.zeros(int(ids.max()) + 1, np.int64); rank_of_id[ids] = rank
static_id = np.zeros(int(ids.max()) + 1, bool); static_id[ids[:3]] = True
print(m.dtype.names)
r, i = m["ref_id"].astype(np.int64), m["inc_id"].astype(np.int64)
rr, ri = rank_of_id[r], rank_of_id[i]
big = 1 << 40
owner = np.minimum(np.where(static_id[r], big, rr), np.where(static_id[i], big, ri))
B = (n + 147) // 148
tile = owner // B
cnt = np.bincount(tile, minlength=148)
print("manifolds per tile: min %d mean %.0f max %d  (max/mean %.2f)" % (cnt.min(), cnt.mean(), cnt.max(), cnt.max() / cnt.mean()))
col = m["color"] if "color" in m.dtype.names else None
if col is not None:
    tasks = np.zeros(148, np.int64)
    for t in range(148):
        c = np.bincount(col[tile == t]); tasks[t] = np.sum((c + 31) // 32)
    print("warp tasks per tile: min %d mean %.1f max %d (max/mean %.2f)" % (tasks.min(), tasks.mean(), tasks.max(), tasks.max() / tasks.mean()))
# bodies touched from a foreign tile
tr, ti = rr // B, ri // B
cross = (~static_id[r]) & (~static_id[i]) & (tr != ti)
print("manifolds across two tiles: %.1f %%" % (100.0 * cross.mean()))
print(np.sort(cnt)[:8], np.sort(cnt)[-8:])
