"""racecheck target (profiles/r02_compute_sanitizer.txt): the per-world kernels only (batch, batch with in-kernel read-back, one small world, global-slot fallback)"""
import os, sys
sys.path.insert(0, '.')
import torch
from resolve2d_b200 import Batch, Solver, scenes
b = Batch(80, 2.0, 4)
for w in range(80): scenes.build_batch_world(b.world(w), w, nx=8, ny=4)
for _ in range(40): b.process(scenes.DT, 4, 4)
n = b.num_bodies()
pin = lambda *shape: torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
out = {"id": None, "pos": pin(n, 2), "angle": pin(n), "momentum": pin(n, 2), "ang_momentum": pin(n), "aabb": None}
for _ in range(3): b.process_read(scenes.DT, 4, 4, out)
s = Solver(2.0, 4); scenes.setup_0_3_many_boxes(s)
for _ in range(60): s.process(scenes.DT, 4, 4)
m = Solver(2.0, 4); scenes.build_mixed(m, 24, 16, n_large=2)
for _ in range(40): m.process(scenes.DT, 4, 4)
print("done", b.stats().n_manifolds, s.stats().n_manifolds, m.stats().n_manifolds)
