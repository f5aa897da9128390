"""Decode SASS control codes (stall counts, barriers) of a kernel's hottest loop.  Usage: sass_loop.py <so> <kernel-substring>"""
import re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = txt.split("Function : ")
blk = [b for b in blocks if pat in b.split("\n")[0]][0]
lines = blk.split("\n")
ins = []
i = 0
while i < len(lines):
    m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s*/\* (0x[0-9a-f]{16}) \*/', lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf, (hi >> 45) & 1, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f))
            i += 2
            continue
    i += 1
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else None
hi_ = int(sys.argv[4], 16) if len(sys.argv) > 4 else None
if lo is None:
    # innermost loop containing the last CREDUX
    cr = [k for k, x in enumerate(ins) if x[1].startswith("CREDUX")]
    k0 = cr[-1]
    for k in range(k0, len(ins)):
        m = re.search(r'BRA (0x[0-9a-f]+)', ins[k][1])
        if m and int(m.group(1), 16) < ins[cr[0]][0]:
            lo, hi_ = int(m.group(1), 16), ins[k][0]
            break
print(f"{len(ins)} instructions; loop {lo:#x}..{hi_:#x} = {(hi_ - lo) // 16 + 1} instructions")
for x in ins:
    if lo <= x[0] <= hi_:
        print(f"{x[0]:05x} st={x[2]:2d} y={x[3]} wb={x[4]} rb={x[5]} wt={x[6]:06b}  {x[1][:80]}")
