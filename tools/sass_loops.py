#!/usr/bin/env python3
"""List the loops (backward branches) of one kernel in a cuobjdump -sass dump with their instruction mix.
usage: cuobjdump -sass lib.so | tools/sass_loops.py <kernel-name-substring> [min-body-instructions]"""
import re, sys
from collections import Counter
name = sys.argv[1]; minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 60
ins = []; on = False
for line in sys.stdin:
    if 'Function :' in line:
        on = name in line
        if on and ins: break
        continue
    if not on: continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(ins))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r'\bBRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\w*\)?', t)
    if 'BRA' in t:
        m2 = re.search(r'0x([0-9a-f]+)', t)
        if not m2: continue
        tgt = int(m2.group(1), 16)
        if tgt < a and tgt in addr_idx:
            j = addr_idx[tgt]; n = i - j + 1
            if n < minlen: continue
            c = Counter()
            for _, tt in ins[j:i + 1]:
                p = tt.split(); op = p[1] if p[0].startswith('@') else p[0]
                c[op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LD', 'ST', 'VI', 'IMAD')) and '.' in op else '')] += 1
            print(f"loop @{tgt:#x}..{a:#x}: {n} instr | " + ", ".join(f"{k} {v}" for k, v in c.most_common(22)))
