#!/usr/bin/env python3
"""Hot loops of a kernel in a cuobjdump -sass dump: size, spill instructions, add-max count.
usage: cuobjdump -sass lib.so | tools/sass_hot.py <kernel-substring> [min] [max]"""
import re, sys
name = sys.argv[1]; lo = int(sys.argv[2]) if len(sys.argv) > 2 else 500; hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1700
ins = []; on = False
for l in sys.stdin:
    if 'Function :' in l:
        if on and ins: break
        on = name in l; continue
    if not on: continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
for i, (a, t) in enumerate(ins):
    if 'BRA' in t:
        m = re.search(r'0x([0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a:
                n = (a - tgt) // 16 + 1
                if lo < n < hi:
                    body = [tt for (ad, tt) in ins if tgt <= ad <= a]
                    print("%#x..%#x %5d instr | spill %3d | VIADDMNMX %4d | LDG %3d | IMAD.MOV %3d" % (
                        tgt, a, n, sum('STL' in x or 'LDL' in x for x in body), sum('VIADDMNMX' in x for x in body),
                        sum('LDG' in x for x in body), sum('IMAD.MOV' in x for x in body)))
