#!/usr/bin/env python
"""Static SASS of one kernel of a built library: tools/dev/sass_fn.py LIB SUBSTR [out.sass] -> instruction count + opcode mix"""
import collections, re, subprocess, sys
lib, sub = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, fns = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1); fns[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        fns[cur].append(line)
for name, lines in fns.items():
    if sub in name:
        ops = collections.Counter()
        for l in lines:
            t = l.split("*/", 1)[1].split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0].rstrip(";")] += 1
        print(name, len(lines))
        print("; ".join(f"{c} {o}" for o, c in ops.most_common(26)))
        if len(sys.argv) > 3:
            open(sys.argv[3], "w").write("\n".join(lines) + "\n")
