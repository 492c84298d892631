"""Strong-scaling probe on ONE GPU: cfg5's 4,096 worlds cut into the per-GPU shares of 1/2/4/8 GPUs
(4096, 2048, 1024, 512 worlds), per-kernel-class device times of each (CUDA events around every launch)."""
import sys, time
sys.path.insert(0, '.')
from resolve2d_b200 import Batch, scenes
sizes = [int(x) for x in sys.argv[1:]] or [4096, 2048, 1024, 512]
for n in sizes:
    b = Batch(n, 2.0, 4)
    for w in range(n): scenes.build_batch_world(b.world(w), w)
    for _ in range(100): b.process(scenes.DT, 4, 4)
    b.reorder()
    for _ in range(5): b.process(scenes.DT, 4, 4)
    b.synchronize(); t = time.perf_counter()
    for _ in range(50): b.process(scenes.DT, 4, 4)
    b.synchronize(); dt = (time.perf_counter() - t) / 50
    b.profile_enable(True)
    for _ in range(10): b.process(scenes.DT, 4, 4)
    p = b.profile_read(); st = b.stats()
    print(f"batch{n}x256: {dt*1e3:.3f} ms/step wall = {n/dt:,.0f} world-steps/s; P={st.n_pairs} M={st.n_manifolds} colours={st.n_colors} launches={st.n_launches}")
    print("   " + "  ".join(f"{k}={ms/10*1e3:.1f}us" for k, (ms, c) in p.items() if c))
    b.destroy()
