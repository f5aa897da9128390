"""Instruction mix of the LAP scan loop(s) of a solve kernel: the innermost backward-branch ranges that contain a warp
reduction (CREDUX).  Usage: sass_scan_loop.py <so> <kernel-substring> [--dump]"""
import collections
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blk = [b for b in txt.split("Function : ") if pat in b.split("\n")[0]][0]
ins = []
for line in blk.split("\n"):
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: k for k, (a, _) in enumerate(ins)}
loops = []
for k, (a, t) in enumerate(ins):
    m = re.search(r'\bBRA(?:\.\w+)* (?:.*, )?(0x[0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr:
            loops.append((addr[tgt], k))
red = [k for k, (_, t) in enumerate(ins) if "CREDUX" in t or "REDUX" in t]
inner = []
for lo, hi in loops:
    if any(lo <= r <= hi for r in red) and not any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi and any(l2 <= r <= h2 for r in red) for l2, h2 in loops):
        inner.append((lo, hi))
print(f"{len(ins)} instructions in kernel; {len(inner)} innermost loop(s) with a warp reduction")
PIPE = {"ALU": ("LOP3", "SEL", "FSEL", "ISETP", "SHF", "IADD3", "IMNMX", "VIMNMX", "PLOP3", "PRMT", "LEA", "VIADD", "IABS", "MOV", "FSETP", "FMNMX", "SGXT", "BMSK", "P2R", "R2P", "CS2R", "POPC", "FLO"),
        "FP64": ("DADD", "DSETP", "DMUL", "DFMA"), "FMA": ("IMAD", "FADD", "FMUL", "FFMA"),
        "XU": ("F2F", "I2F", "F2I", "MUFU"), "LSU": ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDTM", "STTM", "LDC", "ATOM"),
        "REDUX": ("CREDUX", "REDUX", "SHFL", "VOTE", "MATCH", "R2UR", "S2UR"), "CTRL": ("BRA", "BSSY", "BSYNC", "WARPSYNC", "EXIT", "NOP", "CALL", "RET", "YIELD", "ENDCOLLECTIVE", "BREAK")}
for lo, hi in inner:
    ops = collections.Counter()
    pipes = collections.Counter()
    for a, t in ins[lo:hi + 1]:
        t = re.sub(r'^@!?U?P\d+\s+', '', t)
        op = t.split()[0].split(".")[0]
        ops[op] += 1
        pipes[next((p for p, names in PIPE.items() if op in names), "other:" + op)] += 1
    print(f"loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {hi - lo + 1} instructions")
    print("  pipes:", dict(pipes.most_common()))
    print("  ops  :", dict(ops.most_common()))
    if "--dump" in sys.argv:
        for a, t in ins[lo:hi + 1]:
            print(f"    {a:05x}  {t}")
