import json, sys
d = json.loads(open(sys.argv[1]).read())
for k in ("value", "ms_per_step", "gpu_launches", "clocks", "e2e", "cpu_baseline", "batched"):
    print(k, d.get(k))
print(json.dumps(d["config"]["counts"]))
for k, v in d["kernels"].items():
    print(k, {a: round(b, 3) for a, b in v.items()})
