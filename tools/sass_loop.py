#!/usr/bin/env python
"""Instruction count and mix of the innermost hot loop (the smallest backward-branch loop that contains a
global RED and a MUFU) of every kernel matching a name filter in a cuobjdump -sass listing.
Usage: cuobjdump -sass x.o > x.sass; python tools/sass_loop.py x.sass render_bwd"""
import re, sys
from collections import Counter
txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
need = sys.argv[3].split(",") if len(sys.argv) > 3 else ["REDG", "MUFU"]
for f in re.split(r'\n\s+Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if flt not in name:
        continue
    lines = [re.sub(r'/\* 0x[0-9a-f]+ \*/', '', l).rstrip() for l in f.split('\n') if re.match(r'\s+/\*[0-9a-f]{4}\*/', l)]
    addr = lambda l: int(re.match(r'\s*/\*([0-9a-f]{4})\*/', l).group(1), 16)
    best = None
    for l in lines:
        mm = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)', l)
        if mm:
            tgt = int(mm.group(1), 16)
            if tgt < addr(l):
                body = [x for x in lines if tgt <= addr(x) <= addr(l)]
                if all(any(n in x for x in body) for n in need):
                    if best is None or len(body) < best[0]:
                        best = (len(body), body)
    short = re.sub(r'.*?(render_\w+?_kernel\w*?E)v.*', r'\1', name)
    if best is None:
        print(short, "no loop found")
        continue
    op = lambda x: re.sub(r'^@!?U?P\d\s+', '', x.split('*/')[1].strip()).split()[0].split('.')[0]
    c = Counter(op(x) for x in best[1])
    print(short, 'total', len(lines), 'loop', best[0], dict(c.most_common(16)))
