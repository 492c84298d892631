import sys, time
sys.path.insert(0, '.')
from resolve2d_b200 import Solver, scenes
s = Solver(2.0, 4, device=0)
scenes.build_pile100k(s)
for _ in range(200): s.process(scenes.DT, 4, 4)
s.reorder()
for _ in range(5): s.process(scenes.DT, 4, 4)
s.synchronize()
t=time.perf_counter()
for _ in range(50): s.process(scenes.DT, 4, 4)
base=(time.perf_counter()-t)/50
t=time.perf_counter()
for _ in range(10):
    s.reorder(); s.process(scenes.DT, 4, 4)
withr=(time.perf_counter()-t)/10
print(f"process {base*1e3:.3f} ms; reorder+process {withr*1e3:.3f} ms -> reorder costs {1e3*(withr-base):.2f} ms")
s.set_reorder_interval(0)
t=time.perf_counter()
for _ in range(600): s.process(scenes.DT, 4, 4)
print(f"600 steps without any reorder: {(time.perf_counter()-t)/600*1e3:.3f} ms/step")
t=time.perf_counter()
for _ in range(50): s.process(scenes.DT, 4, 4)
print(f"after 650 steps without reorder: {(time.perf_counter()-t)/50*1e3:.3f} ms/step")
