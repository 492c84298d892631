import sys, time
sys.path.insert(0, '.')
from resolve2d_b200 import Solver, Batch, scenes
s = Solver(2.0, 4); cfg = scenes.build_pile100k(s)
t=time.time()
for k in range(1, 501):
    s.process(scenes.DT, 4, 4)
    if k % 50 == 0:
        s.synchronize(); st = s.stats()
        print(k, f"{(time.time()-t)/50*1e3:.2f} ms/step", "E",st.n_entries,"P",st.n_pairs,"M",st.n_manifolds,"K",st.n_points,"col",st.n_colors,"rounds",st.n_color_rounds,"launches",st.n_launches, flush=True)
        t=time.time()
b = Batch(256, 2.0, 4)
for w in range(256): scenes.build_batch_world(b.world(w), w)
t=time.time()
for k in range(1, 301):
    b.process(scenes.DT, 4, 4)
    if k % 50 == 0:
        b.synchronize(); st = b.stats()
        print("batch", k, f"{(time.time()-t)/50*1e3:.2f} ms/step", "P",st.n_pairs,"M",st.n_manifolds,"K",st.n_points,"col",st.n_colors,"rounds",st.n_color_rounds, flush=True)
        t=time.time()
