"""Joins an ncu source-page CSV (SASS view) of one kernel with `nvdisasm -g` line info of the same cubin and prints the
warp-stall samples per CUDA source line.   usage: stalls_by_line.py <src.csv> <dis.txt> <mangled kernel name>"""
import csv, re, sys, collections
src_csv, dis, kname = sys.argv[1:4]
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kname + ':'))
cur = None; sass = []
for l in lines[start + 1:]:
    if l.startswith('//-----'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), 'inlined' in m.group(3)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: sass.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; si = hdr.index('# Samples'); so = hdr.index('Source')
data = [r for r in rows[2:] if len(r) > si]
# ncu rows may be duplicated; keep unique by address
seen = set(); uniq = []
for r in data:
    if r[0] in seen: continue
    seen.add(r[0]); uniq.append(r)
print(len(uniq), "ncu instr rows;", len(sass), "nvdisasm instrs")
agg = collections.Counter(); tot = 0
for r, (txt, loc) in zip(uniq, sass):
    n = int(r[si]) if r[si].isdigit() else 0
    agg[loc[:2] if loc else None] += n; tot += n
srcs = {}
for (loc, n) in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 40):
    if loc is None: print(f"{n:7d} {100*n/tot:5.1f}%  <no line info>"); continue
    f, ln = loc
    if f not in srcs:
        try: srcs[f] = open('resolve2d_b200/csrc/' + f).read().split('\n')
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:110] if 0 < ln <= len(srcs[f]) else ''
    print(f"{n:7d} {100*n/tot:5.1f}%  {f}:{ln:<4d} {text}")
