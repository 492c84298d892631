"""Turns `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table kept under profiles/ and, optionally, into
profiles/traffic.json (DRAM bytes per launch, summed per kernel class).   usage: ncu_table.py raw.csv [workload traffic.json]"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
idx = {c: hdr.index(c) for c in cols if c in hdr}
ki = hdr.index("Kernel Name")
print(" | ".join(["Kernel Name"] + [f"{c}[{units[idx[c]]}]" for c in idx]))
CLASS = [("k_grid_cells|k_scan_chained|k_list_buckets|k_sort_buckets|k_bucket_|k_fine_pairs|k_world_broad", "broadphase"),
         ("k_narrow", "narrowphase"), ("k_color|k_owner|k_partition|k_scan_owners", "coloring"),
         ("k_solve|k_integrate|k_world_solve", "solve_contacts")]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
traffic = {}
scan_seen = 0
for r in data:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("r2d::", "")
    if "at::" in name: continue
    print(" | ".join([name] + [r[idx[c]] for c in idx]))
    b = sum(to_bytes(r[idx[c]], units[idx[c]]) for c in ("dram__bytes_read.sum", "dram__bytes_write.sum") if c in idx)
    cls = next((c for pat, c in CLASS if re.search(pat, name)), None)
    if cls: traffic[cls] = traffic.get(cls, 0.0) + b
if len(sys.argv) > 3:
    path = sys.argv[3]
    try: t = json.load(open(path))
    except Exception: t = {}
    t["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per process() call, summed per kernel class, from one "
                  "`ncu --set full` capture of one step (profiles/ncu_table.py)")
    t[sys.argv[2]] = traffic
    json.dump(t, open(path, "w"), indent=1)
