import sys, time
sys.path.insert(0, '.')
from resolve2d_b200 import Solver, scenes
s = Solver(2.0, 4); scenes.build_pile100k(s)
for _ in range(50): s.process(scenes.DT, 4, 4)
s.synchronize()
for k in range(4):
    t = time.perf_counter(); s.reorder(); s.synchronize(); t1 = time.perf_counter()
    s.process(scenes.DT, 4, 4); s.synchronize(); t2 = time.perf_counter()
    print(f"reorder {1e3*(t1-t):.3f} ms, next process {1e3*(t2-t1):.3f} ms")
