"""Static SASS statistics of one kernel from `nvdisasm -g -c` output (built with -lineinfo): instruction count, and the
source lines that own the spill instructions (STL/LDL) and the most code.

  cuobjdump -xelf all rt_trace.o && nvdisasm -g -c rt_trace.sm_100a.cubin > all.sass
  python tools/sass_lines.py all.sass <mangled-name-substring> [top]
"""
import re
import sys
from collections import Counter

lines = open(sys.argv[1]).read().split("\n")
key = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and key in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.") or lines[i].startswith("\t.section")), len(lines))
cur, code, spill, ops = None, Counter(), Counter(), Counter()
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m:
        code[cur] += 1
        ops[m.group(1)] += 1
        if m.group(1) in ("STL", "LDL"):
            spill[cur] += 1
print("instructions:", sum(code.values()), " spill instructions:", sum(spill.values()))
print("top opcodes:", ops.most_common(14))
print("spills by line:", sorted(spill.items(), key=lambda x: -x[1])[:top])
print("code by line:", sorted(code.items(), key=lambda x: -x[1])[:top])
