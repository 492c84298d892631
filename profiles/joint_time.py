import os, sys, time
sys.path.insert(0, '.')
from resolve2d_b200 import Batch, Solver, scenes
def t_single(build, S, I):
    s = Solver(2.0, 4); build(s)
    for _ in range(60): s.process(scenes.DT, S, I)
    s.synchronize(); t = time.perf_counter()
    for _ in range(100): s.process(scenes.DT, S, I)
    s.synchronize(); dt = (time.perf_counter() - t) * 10
    return dt, s.stats().n_launches
def t_batch(n):
    b = Batch(n, 2.0, 4)
    for w in range(n): scenes.build_pyramid(b.world(w), base=6 + w % 3, n_spinners=1 + w % 2)
    for _ in range(40): b.process(scenes.DT, 4, 10)
    b.synchronize(); t = time.perf_counter()
    for _ in range(50): b.process(scenes.DT, 4, 10)
    b.synchronize(); dt = (time.perf_counter() - t) * 20
    return dt, b.stats().n_launches
for flag in ("1", "0"):
    os.environ["R2D_WORLD_JOINTS"] = flag
    print("R2D_WORLD_JOINTS=" + flag,
          "0_1 car platformer: %.3f ms (%d launches)" % t_single(scenes.setup_0_1_car_platformer, 4, 4),
          "| pyramid40: %.3f ms (%d launches)" % t_single(lambda s: scenes.build_pyramid(s, base=40, n_spinners=6), 4, 10),
          "| 512 jointed worlds: %.3f ms (%d launches)" % t_batch(512))
