"""Per-kernel-class device time of process() on a workload (CUDA events around every launch)."""
import sys, time
sys.path.insert(0, '.')
from resolve2d_b200 import Solver, scenes
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "pile100k"
build, preroll = bench.workload_table()[name]
import os
s = Solver(float(os.environ.get("R2D_CELL", "2.0")), int(os.environ.get("R2D_MULT", "4")))
if os.environ.get("R2D_FAST"): s.set_mode(1)
cfg = build(s)
S, I = cfg["sub_steps"], cfg["iters"]
for _ in range(preroll + 50): s.process(scenes.DT, S, I)
s.synchronize(); t = time.perf_counter()
for _ in range(100): s.process(scenes.DT, S, I)
s.synchronize(); dt = (time.perf_counter() - t) / 100
s.profile_enable(True)
for _ in range(20): s.process(scenes.DT, S, I)
p = s.profile_read(); st = s.stats()
print(f"{name}: {dt*1e3:.3f} ms/step wall; N={st.n_bodies} E={st.n_entries} P={st.n_pairs} M={st.n_manifolds} colours={st.n_colors} rounds={st.n_color_rounds} launches={st.n_launches}")
for k, (ms, n) in p.items():
    if n: print(f"   {k:15s} {ms/20*1e3:8.1f} us/step  {n/20:5.1f} launches/step")
