"""Instruction mix of the loops of one kernel: python tools/sass_loops.py lib.so 'gp_chunk_kernelILi3ELb1' [min_dfma]"""
import re
import subprocess
import sys
from collections import Counter

lib, pat = sys.argv[1], sys.argv[2]
min_dfma = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = txt.split("Function : ")
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in f.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(name, len(ins), "instructions")
    for i, (addr, text) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= addr:
            continue
        body = [t for a, t in ins if tgt <= a <= addr]
        ops = Counter()
        for t in body:
            t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
            ops[t2.split()[0].split(".")[0]] += 1
        if ops["DFMA"] >= min_dfma:
            fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "F2F", "FRND", "F2I", "I2F", "DMNMX"))
            print(f"loop {tgt:#x}..{addr:#x}: {len(body)} instr, FP64-pipe-ish {fp64}, BRA inside {ops['BRA']}, mix {dict(ops.most_common(14))}")
