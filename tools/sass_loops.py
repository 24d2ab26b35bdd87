"""Development aid: loops of one kernel's SASS (cuobjdump -sass -fun <mangled> lib.so > f.sass) with their size, loads, stores,
FP64 ops and local-memory (spill) traffic.  usage: sass_loops.py f.sass"""
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr = [a for a, _ in ins]
loops = []
for a, t in ins:
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a:
            loops.append((tgt, a))
loops.sort(key=lambda l: l[1] - l[0])
print(f"{len(ins)} instructions, {len(loops)} backward branches")
for lo, hi in loops:
    body = [t for a, t in ins if lo <= a <= hi]
    inner = [l for l in loops if l != (lo, hi) and lo <= l[0] and l[1] <= hi]
    cnt = lambda pat: sum(1 for t in body if re.search(pat, t))
    print(f"loop {lo:#07x}-{hi:#07x}: {len(body):5d} instr, LDG {cnt(r'LDG'):3d} STG {cnt(r'STG'):3d} DFMA/DMUL/DADD {cnt(r'DMUL|DADD|DFMA'):3d} "
          f"LDL {cnt(r'LDL'):3d} STL {cnt(r'STL'):3d} CCTL/prefetch {cnt(r'CCTL'):3d} inner loops {len(inner)}")
