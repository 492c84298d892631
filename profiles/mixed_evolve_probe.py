import sys, time
sys.path.insert(0, '.')
from resolve2d_b200 import Solver, scenes
for name, nx, ny in (("round 1: 5000 x 200", 5000, 200), ("round 2: 20000 x 50", 20000, 50)):
    s = Solver(2.0, 4); scenes.build_mixed(s, nx, ny, 1000)
    t = time.time()
    for k in range(1, 301):
        s.process(scenes.DT, 4, 4)
        if k in (20, 50, 90, 100, 150, 200, 250, 300):
            s.synchronize(); st = s.stats()
            print(f"| {name} | {k} | {st.n_pairs} | {st.n_manifolds} | {st.n_colors} | {(time.time()-t)*1e3/ (k if k==20 else 1):.0f} |", flush=True)
    s.deinit()
