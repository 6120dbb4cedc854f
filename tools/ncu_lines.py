#!/usr/bin/env python
"""Dynamic instruction counts per source line: joins `ncu --page source --csv` (SASS view: address, executed count) with
`nvdisasm -g -c` of the cubin that ran (built with -lineinfo), matching instructions by position inside each function.

  cuobjdump -xelf all librt_b200.so && nvdisasm -g -c rt_trace.sm_100a.cubin > rt_trace.sass
  ncu -i X.ncu-rep --page source --csv > X_src.csv
  python tools/ncu_lines.py X_src.csv rt_trace.sass <kernel-mangled-substring> [top]
"""
import csv
import re
import sys
from collections import Counter, defaultdict

src_csv, sass, key = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# ---- nvdisasm: per function, list of (offset, opcode, file, line)
ANCHOR = "rt_trace.cu"      # with `nvdisasm -gi`, inlined helper code is charged to the statement of this file that called it
funcs = {}
cur_fn, cur_line = None, None
chain, pending = [], False
for l in open(sass):
    if l.startswith(".text.") or l.startswith("\t.section\t.text."):
        cur_fn = l.strip().split(".text.")[1].split(",")[0].rstrip(":")
        funcs[cur_fn] = []
        cur_line = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        fr = (m.group(1).split("/")[-1], int(m.group(2)))
        if not pending:
            chain = []
        pending = True
        if not chain or chain[-1] != fr:
            chain.append(fr)
        if m.group(3):
            chain.append((m.group(3).split("/")[-1], int(m.group(4))))
        # attribute to the innermost frame inside the anchor file (nvdisasm -gi), else to the innermost frame
        anchored = [c for c in chain if c[0] == ANCHOR]
        cur_line = anchored[0] if anchored else chain[0]
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        pending = False
    if m and cur_fn:
        funcs[cur_fn].append((int(m.group(1), 16), m.group(2), cur_line))

# ---- ncu rows
rows = []
with open(src_csv) as f:
    rd = csv.reader(f)
    hdr = None
    for r in rd:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
            rows.append((int(r[0], 16), r[1].strip(), int(r[hdr.index("Instructions Executed")] or 0),
                         int(r[hdr.index("# Samples")] or 0)))
rows.sort()
# split ncu rows into contiguous address runs (one per function)
runs, cur = [], [rows[0]]
for a in rows[1:]:
    if a[0] - cur[-1][0] == 16:
        cur.append(a)
    else:
        runs.append(cur); cur = [a]
runs.append(cur)
print("ncu: %d instructions in %d address runs: %s" % (len(rows), len(runs), [len(r) for r in runs]))

def opcode(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0].rstrip(";")

by_line, by_fn, samples = Counter(), Counter(), Counter()
total = 0
for run in runs:
    # find the function with the same length and opcode sequence
    ops = [opcode(r[1]) for r in run]
    cands = [fn for fn, ins in funcs.items() if len(ins) == len(run) and [i[1] for i in ins] == ops]
    cands = [c for c in cands if key in c] or cands
    if not cands:
        cands = [fn for fn, ins in funcs.items() if len(ins) == len(run)]
    if not cands:
        print("  run of %d instructions: no matching function" % len(run)); continue
    fn = cands[0]
    for (addr, text, n, smp), (off, op, line) in zip(run, funcs[fn]):
        by_line[line] += n; by_fn[fn[-60:]] += n; samples[line] += smp; total += n
print("total warp instructions: %.4g" % total)
for fn, n in by_fn.most_common():
    print("  %5.1f %%  %s" % (100.0 * n / total, fn))
print("by line (share of executed warp instructions, share of stall samples):")
ts = sum(samples.values()) or 1
for line, n in by_line.most_common(top):
    print("  %5.2f %%  %5.2f %%  %s" % (100.0 * n / total, 100.0 * samples[line] / ts, line))
if len(sys.argv) > 5:       # dump everything for bucketing
    with open(sys.argv[5], "w") as f:
        for line, n in sorted(by_line.items(), key=lambda x: (str(x[0]))):
            f.write("%s\t%s\t%d\t%d\n" % (line[0] if line else None, line[1] if line else 0, n, samples[line]))
