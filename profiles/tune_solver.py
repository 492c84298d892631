"""Times process() on pile100k for several wait-policy settings of the dataflow sweep (env-tunable, results identical)."""
import os, sys, time, itertools
sys.path.insert(0, '.')
def run(env):
    for k, v in env.items(): os.environ[k] = str(v)
    from resolve2d_b200 import Solver, scenes
    s = Solver(2.0, 4); scenes.build_pile100k(s)
    for _ in range(250): s.process(scenes.DT, 4, 4)
    s.synchronize(); t = time.perf_counter()
    for _ in range(100): s.process(scenes.DT, 4, 4)
    s.synchronize(); dt = (time.perf_counter() - t) / 100
    s.profile_enable(True)
    for _ in range(20): s.process(scenes.DT, 4, 4)
    p = s.profile_read()
    print(env, f"{dt*1e3:.3f} ms/step  solver {p['solve_contacts'][0]/20*1e3:.0f} us", flush=True)
    s.deinit()
if __name__ == "__main__":
    import subprocess, json
    if len(sys.argv) > 1:
        run(json.loads(sys.argv[1]))
    else:
        cfgs = []
        for b, mode in ((1, 3), (4, 3), (4, 2)):
            for spin, unit, mx in ((1, 200, 4000),):
                cfgs.append(dict(R2D_SOLVE_BLOCKS_PER_SM=b, R2D_WAIT_MODE=mode, R2D_WAIT_SPIN_LAG=spin, R2D_WAIT_SLEEP_UNIT=unit, R2D_WAIT_SLEEP_MAX=mx))
        for c in cfgs:
            subprocess.run([sys.executable, __file__, json.dumps(c)])
