import os, sys
sys.path.insert(0, '.')
from resolve2d_b200 import Solver, scenes
s = Solver(2.0, 4); scenes.build_pile100k(s)
for _ in range(250): s.process(scenes.DT, 4, 4)
os.environ["R2D_STAMPS"] = "1"
for _ in range(3): s.process(scenes.DT, 4, 4)
