"""Joins an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --kernel-name K) with `nvdisasm -g -c` of the
same cubin and prints, per CUDA source line: warp instructions executed, average active lanes, share of stall samples.
usage: lines_by_metric.py <src.csv> <dis.txt> <kernel substring> [top N]"""
import csv, re, sys, collections
src_csv, dis, ksub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split('\n')
kname = [l for l in lines if l.startswith('.text.') and ksub in l][0][6:-1]
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kname + ':'))
cur = None; sass = []
for l in lines[start + 1:]:
    if l.startswith('//-----'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: sass.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ie = hdr.index('Instructions Executed'); it = hdr.index('Thread Instructions Executed'); isamp = hdr.index('# Samples')
seen = set(); uniq = []
for r in rows[2:]:
    if len(r) <= ie or r[0] in seen: continue
    seen.add(r[0]); uniq.append(r)
print(kname, len(uniq), "ncu rows;", len(sass), "nvdisasm instrs")
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
for r, (txt, loc) in zip(uniq, sass):
    e = int(r[ie]); t = int(r[it]); sm = int(r[isamp]) if r[isamp].isdigit() else 0
    a = agg[loc]; a[0] += e; a[1] += t; a[2] += sm; tot[0] += e; tot[1] += t; tot[2] += sm
print("total warp-inst", tot[0], "avg lanes %.1f" % (tot[1] / max(tot[0], 1)), "samples", tot[2])
srcs = {}
key = (lambda kv: -kv[1][2]) if (len(sys.argv) > 5 and sys.argv[5] == 'samples') else (lambda kv: -kv[1][0])
for loc, a in sorted(agg.items(), key=key)[:top]:
    f, ln = loc if loc else ('?', 0)
    if f not in srcs:
        try: srcs[f] = open('resolve2d_b200/csrc/' + f).read().split('\n')
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ''
    print(f"{a[0]:9d} {100*a[0]/tot[0]:5.1f}% lanes {a[1]/max(a[0],1):5.1f} samp {100*a[2]/max(tot[2],1):5.1f}%  {f}:{ln:<4d} {text}")
