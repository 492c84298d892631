"""Per-kernel-class device time of batch_process() on cfg5 (n worlds x 256 bodies)."""
import sys, time
sys.path.insert(0, '.')
from resolve2d_b200 import Batch, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = Batch(n, 2.0, 4)
for w in range(n): scenes.build_batch_world(b.world(w), w)
for _ in range(100): b.process(scenes.DT, 4, 4)
b.reorder()
for _ in range(5): b.process(scenes.DT, 4, 4)
b.synchronize(); t = time.perf_counter()
for _ in range(30): b.process(scenes.DT, 4, 4)
b.synchronize(); dt = (time.perf_counter() - t) / 30
b.profile_enable(True)
for _ in range(10): b.process(scenes.DT, 4, 4)
p = b.profile_read(); st = b.stats()
print(f"batch{n}x256: {dt*1e3:.3f} ms/step wall = {n/dt:,.0f} world-steps/s; P={st.n_pairs} M={st.n_manifolds} colours={st.n_colors} rounds={st.n_color_rounds} launches={st.n_launches}")
for k, (ms, c) in p.items():
    if c: print(f"   {k:15s} {ms/10*1e3:8.1f} us/step  {c/10:5.1f} launches/step")
