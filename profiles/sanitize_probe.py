"""Small workloads for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the pipeline runs, on a
single world with joints and exclusions, a mixed world with multi-cell bodies, and a batch large enough for the
CTA-per-world solver."""
import sys
sys.path.insert(0, '.')
from resolve2d_b200 import Batch, Solver, scenes
s = Solver(2.0, 4); scenes.setup_0_1_car_platformer(s)
for _ in range(110): scenes.drive_0_1(s); s.process(scenes.DT, 4, 4)
s.read_bodies(); s.read_pairs(); s.read_manifolds()
m = Solver(2.0, 4); scenes.build_mixed(m, 60, 30, n_large=4)
for _ in range(70): m.process(scenes.DT, 4, 4)
m.reorder(); m.process(scenes.DT, 4, 4); m.read_bodies()
b = Batch(80, 2.0, 4)
for w in range(80): scenes.build_batch_world(b.world(w), w, nx=8, ny=4)
for _ in range(80): b.process(scenes.DT, 4, 4)
b.read_bodies()
print("sanitize probe done", s.stats().n_manifolds, m.stats().n_manifolds, b.stats().n_manifolds)
