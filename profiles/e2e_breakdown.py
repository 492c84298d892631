"""Host wall-clock split of the end-to-end step (pinned host buffers in and out through the C ABI) on pile100k."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from resolve2d_b200 import Solver, scenes
s = Solver(2.0, 4, device=0)
cfg = scenes.build_pile100k(s)
S, I, dt = cfg["sub_steps"], cfg["iters"], scenes.DT
for _ in range(200): s.process(dt, S, I)
s.reorder()
n = s.num_bodies()
forces = torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy()
out = {k: torch.empty(shape, dtype=torch.float32).pin_memory().numpy() for k, shape in
       (("pos", (n, 2)), ("angle", (n,)), ("momentum", (n, 2)), ("ang_momentum", (n,)))}
out["id"] = None; out["aabb"] = None
for _ in range(5):
    s.write_forces(forces); s.process(dt, S, I); s.read_bodies(out)
K = 100
t = [0.0, 0.0, 0.0]
t0 = time.perf_counter()
for _ in range(K):
    a = time.perf_counter(); s.write_forces(forces)
    b = time.perf_counter(); s.process(dt, S, I)
    c = time.perf_counter(); s.read_bodies(out)
    e = time.perf_counter()
    t[0] += b - a; t[1] += c - b; t[2] += e - c
tot = time.perf_counter() - t0
print(f"e2e {tot/K*1e6:.1f} us/step: write_forces {t[0]/K*1e6:.1f}  process {t[1]/K*1e6:.1f}  read_bodies {t[2]/K*1e6:.1f}")
t0 = time.perf_counter()
for _ in range(K): s.process(dt, S, I)
print(f"process only (warm L2, host wall) {(time.perf_counter()-t0)/K*1e6:.1f} us/step")
