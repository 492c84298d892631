"""Transcribes the golden vectors of SURVEY.md Appendix F (per-step hashes and raw f32 bit patterns that the survey
obtained by executing the reference's own prebuilt binary, /root/reference/demos/web/public/resolve2d.wasm) into
tests/golden/appendix_f.json.  Run from the repo root:  python tests/golden/make_appendix_f.py
(tests/golden/make_wasm_golden.py re-derives the state/aabb columns here by running that wasm again.)"""
import json
import re

text = open("SURVEY.md").read()
sections = {"0_1_car_platformer": "### F.2", "0_1_car_platformer_driven": "### F.3", "0_3_many_boxes": "### F.4"}
ends = {"### F.2": "### F.3", "### F.3": "### F.4", "### F.4": "### F.5"}
out = {}
for name, start in sections.items():
    seg = text[text.index(start):text.index(ends[start])]
    m0 = re.search(r"Step 0: state `([0-9a-f]{16})`, aabb `([0-9a-f]{16})`", seg)
    entry = {"steps": [], "raw": []}
    if m0:
        entry["step0"] = {"state": m0.group(1), "aabb": m0.group(2)}
    for line in seg.splitlines():
        m = re.match(r"\| (\d+) \| ([0-9a-f]{16}) \| ([0-9a-f]{16}) \| (\d+) \| (\d+) / ([0-9a-f]{16}) \| (\d+), (\d+) / ([0-9a-f]{16}) \|", line)
        if m:
            entry["steps"].append({"step": int(m.group(1)), "state": m.group(2), "aabb": m.group(3), "E": int(m.group(4)),
                                   "C": int(m.group(5)), "pairs": m.group(6), "M": int(m.group(7)), "K": int(m.group(8)),
                                   "man": m.group(9)})
            continue
        m = re.match(r"\| (\d+) \| (\d+) \| ([0-9a-f]{8}),([0-9a-f]{8}) \| ([0-9a-f]{8}) \| ([0-9a-f]{8}),([0-9a-f]{8}) \| ([0-9a-f]{8}) \| ([0-9a-f]{8}),([0-9a-f]{8}),([0-9a-f]{8}),([0-9a-f]{8}) \|", line)
        if m:
            g = m.groups()
            entry["raw"].append({"step": int(g[0]), "id": int(g[1]), "pos": [g[2], g[3]], "angle": g[4],
                                 "momentum": [g[5], g[6]], "ang_momentum": g[7], "aabb": list(g[8:12])})
    out[name] = entry
out["0_1_car_platformer_driven"]["step0"] = out["0_1_car_platformer"]["step0"]
json.dump(out, open("tests/golden/appendix_f.json", "w"), indent=1)
print({k: (len(v["steps"]), len(v["raw"])) for k, v in out.items()})
