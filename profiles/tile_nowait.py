"""pile100k per-class device times; with R2D_WAIT_MODE=2 the tile solver ignores all dependencies (WRONG results): what is
left is its dependency-free cost (records, arithmetic, body phases, grid barriers)."""
import sys
sys.path.insert(0, '.')
from resolve2d_b200 import Solver, scenes
s = Solver(2.0, 4, device=0)
scenes.build_pile100k(s)
for _ in range(200): s.process(scenes.DT, 4, 4)
s.reorder()
for _ in range(3): s.process(scenes.DT, 4, 4)
s.profile_enable(True)
for _ in range(10): s.process(scenes.DT, 4, 4)
p = s.profile_read()
print({k: round(ms / 10 * 1e3, 1) for k, (ms, c) in p.items() if c})
