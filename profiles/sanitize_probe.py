"""Small workloads for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the pipeline runs, on a
single world with joints and exclusions, a mixed world with multi-cell bodies, and a batch large enough for the
CTA-per-world kernels (k_world_broad, k_world_solve)."""
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
# fine grid with DYNAMIC large bodies: small-large pairs through the coarse buckets, large-large pairs from the bucket kernels
c = Solver(2.0, 4); scenes.build_mixed(c, 40, 12, n_large=6)
for _ in range(230): c.process(scenes.DT, 4, 4)
c.read_pairs()
# batch with joints: per-world sorted colouring + exported masks + persistent dataflow sweep
j = Batch(80, 2.0, 4)
for w in range(80): scenes.build_pyramid(j.world(w), base=6, n_spinners=1)
for _ in range(20): j.process(scenes.DT, 4, 10)
# bulk force import on the side stream, zero-copy export into pinned memory
import numpy as np, torch
n = m.num_bodies()
f = torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy()
out = {"pos": torch.empty((n, 2), dtype=torch.float32).pin_memory().numpy(), "angle": None, "momentum": None, "ang_momentum": None, "id": None, "aabb": None}
for _ in range(3):
    m.write_forces(f); m.process(scenes.DT, 4, 4); m.read_bodies(out)
# roadmap options: warm-start table (k_warm_save, lookup in the pre-step), sleeping kernels; reference-order validation mode
from resolve2d_b200 import MODE_REFERENCE_ORDER, OPT_SLEEP_CALLS, OPT_SLEEPING, OPT_WARM_START, ShardedBatch
o = Solver(2.0, 4); o.set_option(OPT_WARM_START, 1); o.set_option(OPT_SLEEPING, 1); o.set_option(OPT_SLEEP_CALLS, 5)
scenes.build_box1k(o)
for _ in range(100): o.process(scenes.DT, 4, 4)
r = Solver(2.0, 4); r.set_mode(MODE_REFERENCE_ORDER); scenes.setup_0_1_car_platformer(r)
for _ in range(30): r.process(scenes.DT, 4, 4)
h = Solver(2.0, 4); scenes.build_hub(h, n_discs=300)          # colours exhausted on one body: dropped manifolds
for _ in range(10): h.process(scenes.DT, 4, 4)
sb = ShardedBatch(90, [0, 0])                                   # two shards on one device, one host thread each
for w in range(90): scenes.build_batch_world(sb.world(w), w, nx=8, ny=4)
for _ in range(30): sb.process(scenes.DT, 4, 4)
sb.read_bodies()
# device-side re-sort (k_resort_*, cub radix sort) with joints / exclusions / a batch; in-kernel read-back of k_world_solve
s.reorder(); s.process(scenes.DT, 4, 4); j.reorder(); j.process(scenes.DT, 4, 10); b.reorder()
nb_ = b.num_bodies()
pin = lambda *shape: torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
outb = {"id": None, "pos": pin(nb_, 2), "angle": pin(nb_), "momentum": pin(nb_, 2), "ang_momentum": pin(nb_), "aabb": None}
for _ in range(3): b.process_read(scenes.DT, 4, 4, outb)
# list flavour of the dataflow colouring with chained hub lists (normally worlds of ~10^6 bodies)
import os
os.environ["R2D_FLOW_LIST"] = "1"
hl = Solver(2.0, 4); scenes.build_hub(hl)
for _ in range(10): hl.process(scenes.DT, 4, 4)
ml = Solver(2.0, 4); scenes.build_mixed(ml, 40, 20, n_large=3)
for _ in range(40): ml.process(scenes.DT, 4, 4)
os.environ["R2D_SOLVE_WIDE"] = "1"                              # 512-thread persistent sweep + record prefetch
pw = Solver(2.0, 4); scenes.build_pyramid(pw, base=12, n_spinners=1)
for _ in range(10): pw.process(scenes.DT, 4, 10)
print("sanitize probe done", s.stats().n_manifolds, m.stats().n_manifolds, b.stats().n_manifolds)
