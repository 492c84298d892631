"""<scene> [pre-roll steps]: per-class device times and a checksum of the state after 230 calls (flavour switches via env)."""
import sys, time, hashlib
sys.path.insert(0, '.')
import numpy as np
from resolve2d_b200 import Solver, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "pile100k"
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 200
s = Solver(2.0, 4) if name != "mixed1M" else Solver(4.0, 4)
cfg = getattr(scenes, "build_" + name)(s)
S, I = cfg["sub_steps"], cfg["iters"]
for _ in range(pre): s.process(scenes.DT, S, I)
s.reorder()
for _ in range(10): s.process(scenes.DT, S, I)
s.synchronize(); t = time.perf_counter()
for _ in range(50): s.process(scenes.DT, S, I)
s.synchronize(); dt = (time.perf_counter() - t) / 50
s.profile_enable(True)
for _ in range(20): s.process(scenes.DT, S, I)
p = s.profile_read(); st = s.stats()
b = s.read_bodies()
h = hashlib.sha1(np.ascontiguousarray(b["pos"]).tobytes() + np.ascontiguousarray(b["momentum"]).tobytes()).hexdigest()[:12]
print(f"{name}: {dt*1e3:.3f} ms/step wall; P={st.n_pairs} M={st.n_manifolds} colours={st.n_colors} launches={st.n_launches} sha={h}")
print("   " + "  ".join(f"{k}={ms/20*1e3:.1f}us" for k, (ms, c) in p.items() if c))
