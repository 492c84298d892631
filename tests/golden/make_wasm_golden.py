"""Generates tests/golden/wasm_golden.json by EXECUTING THE REFERENCE ITSELF: the prebuilt resolve2d build the
reference ships (/root/reference/demos/web/public/resolve2d.wasm) is run in oracle/wasm_interp.cpp (a small wasm
interpreter; the module has no imports) through its own exports.  For every run: FNV hashes of body state
and of the AABBs after EVERY process() call, plus complete raw body dumps (f32 bit patterns) at a few steps.

Run from the repo root in the build container (the wasm file does not exist on the GPU box; the JSON travels):
    make -C oracle _build/wasm_run && python tests/golden/make_wasm_golden.py
"""
import json
import subprocess
import sys

WASM = "/root/reference/demos/web/public/resolve2d.wasm"
RUNS = {
    # name: (scene, steps, rate (dt = 1/rate in f32), sub_steps, iters, extra args)
    "0_3_many_boxes": ("0_3", 240, 60, 4, 4, ["dump:1,60,120,240"]),
    "0_1_car_platformer": ("0_1", 300, 60, 4, 4, ["dump:1,60,300"]),
    "0_1_car_platformer_driven": ("0_1", 240, 60, 4, 4, ["driven", "dump:240"]),
    "0_1_rate120_s2_i6": ("0_1", 200, 120, 2, 6, ["dump:200"]),
    "0_3_s1_i1": ("0_3", 150, 60, 1, 1, ["dump:150"]),
    "0_3_remove_bodies": ("0_3", 150, 60, 4, 4, ["remove:30:5,100,261", "remove:90:1", "dump:31,150"]),
}
out = {"_source": "resolve2d.wasm (sha256 d7e1ad69...cf91) executed by oracle/wasm_interp.cpp; see make_wasm_golden.py"}
for name, (scene, steps, rate, S, I, extra) in RUNS.items():
    cmd = ["oracle/_build/wasm_run", WASM, scene, str(steps), str(rate), str(S), str(I)] + extra
    print(" ".join(cmd), file=sys.stderr, flush=True)
    res = json.loads(subprocess.check_output(cmd))
    res["removals"] = [a for a in extra if a.startswith("remove:")]
    out[name] = res
json.dump(out, open("tests/golden/wasm_golden.json", "w"), separators=(",", ":"))
print({k: len(v["steps"]) for k, v in out.items() if k != "_source"})
