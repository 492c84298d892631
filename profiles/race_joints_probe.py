"""racecheck target (profiles/r02_compute_sanitizer.txt): joints inside k_world_solve — one small world with all joint kinds and
exclusions, a batch of jointed worlds."""
import sys
sys.path.insert(0, '.')
from resolve2d_b200 import Batch, Solver, scenes
s = Solver(2.0, 4); scenes.setup_0_1_car_platformer(s)
for _ in range(12): scenes.drive_0_1(s); s.process(scenes.DT, 4, 4)
j = Batch(80, 2.0, 4)
for w in range(80): scenes.build_pyramid(j.world(w), base=5, n_spinners=1)
for _ in range(6): j.process(scenes.DT, 4, 6)
print("done", s.stats().n_joints, j.stats().n_joints, j.stats().n_launches)
