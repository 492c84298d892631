"""Colour counts of greedy edge colourings of the contact graph under different orders (CPU, oracle state)."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from oracle import OracleSolver, ORDER_COLORED
from resolve2d_b200 import scenes
which = sys.argv[1] if len(sys.argv) > 1 else "pile"
o = OracleSolver(2.0, 4, order=ORDER_COLORED)
if which == "pile": scenes.build_pile(o, 200, 50); steps = 200
elif which == "batch": scenes.build_batch_world(o, 3); steps = 150
else: scenes.build_box1k(o); steps = 150
t = time.time()
for _ in range(steps): o.process(scenes.DT, 4, 4)
m = o.read_manifolds(); b = o.read_bodies()
print(which, "steps", steps, "%.1fs" % (time.time() - t), "manifolds", len(m), "colours now", int(m["color"].max()) + 1)
ids = b["id"]; nstat = {"pile": 3, "batch": None, "box": None}[which]
# static flags: infer from momentum == 0 and id small?  use oracle API if present
try:
    static = {int(i) for i in ids if o.body_handle(int(i)).is_static()}
except Exception:
    static = set(int(i) for i in ids[:3])
pos = {int(i): p for i, p in zip(ids, b["pos"])}
E = [(int(r), int(i)) for r, i in zip(m["ref_id"], m["inc_id"])]
deg = {}
for a, c in E:
    for x in (a, c):
        if x not in static: deg[x] = deg.get(x, 0) + 1
print("max body degree", max(deg.values()), "mean", np.mean(list(deg.values())))
def greedy(order):
    used = {}
    ncol = 0
    for k in order:
        a, c = E[k]
        u = 0
        if a not in static: u |= used.get(a, 0)
        if c not in static: u |= used.get(c, 0)
        col = 0
        while u >> col & 1: col += 1
        if a not in static: used[a] = used.get(a, 0) | 1 << col
        if c not in static: used[c] = used.get(c, 0) | 1 << col
        ncol = max(ncol, col + 1)
    return ncol
n = len(E)
rng = np.random.default_rng(1)
print("random order:", [greedy(rng.permutation(n)) for _ in range(3)])
dsum = np.array([deg.get(a, 0) + deg.get(c, 0) for a, c in E]); dmax = np.array([max(deg.get(a, 0), deg.get(c, 0)) for a, c in E])
tie = rng.random(n)
print("largest degree-sum first:", greedy(np.lexsort((tie, -dsum))))
print("largest max-degree first:", greedy(np.lexsort((tie, -dmax))))
mid = np.array([(pos[a] + pos[c]) / 2 if (a not in static and c not in static) else (pos[c] if a in static else pos[a]) for a, c in E])
print("by x then y of the contact:", greedy(np.lexsort((mid[:, 1], mid[:, 0]))))
print("by y then x of the contact:", greedy(np.lexsort((mid[:, 0], mid[:, 1]))))
ang = np.array([np.arctan2(*(pos[c] - pos[a])[::-1]) if (a not in static and c not in static) else 9.0 for a, c in E])
# by direction class of the contact (6 sectors of 30 degrees, mod 180), then position
sector = np.where(ang > 8, 6, np.floor(((ang % np.pi) / np.pi) * 6).astype(int) % 6)
print("by direction sector then random:", greedy(np.lexsort((tie, sector))))
print("by direction sector then x,y:", greedy(np.lexsort((mid[:, 1], mid[:, 0], sector))))
