import sys
sys.path.insert(0, '.')
from resolve2d_b200 import Solver, scenes
s = Solver(2.0, 4)
cfg = scenes.setup_0_3_many_boxes(s) or {}
last = None
for i in range(240):
    try:
        s.process(scenes.DT, 4, 4)
    except Exception as e:
        print("step", i, "ERR", e); break
    st = s.stats()
    cur = (st.n_pairs, st.n_manifolds, st.n_colors, st.n_color_rounds)
    if cur != last: print(i, *cur)
    last = cur
