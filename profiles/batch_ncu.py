"""One batch_process() of cfg5 inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`)."""
import ctypes, sys
sys.path.insert(0, '.')
from resolve2d_b200 import Batch, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = Batch(n, 2.0, 4)
for w in range(n): scenes.build_batch_world(b.world(w), w)
for _ in range(100): b.process(scenes.DT, 4, 4)
b.reorder()
for _ in range(3): b.process(scenes.DT, 4, 4)
b.synchronize()
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaProfilerStart()
for _ in range(1): b.process(scenes.DT, 4, 4)
b.synchronize()
rt.cudaProfilerStop()
