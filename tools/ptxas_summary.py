"""Summarises `nvcc -Xptxas -v` output: one line per kernel with registers, spills, stack, shared memory."""
import re
import subprocess
import sys

txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers[^\n]*?(?:(\d+) bytes smem)?\n")
for m in pat.finditer(txt):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = name.replace("(anonymous namespace)::", "").replace("void ", "")
    print("%-72s regs %3s  stack %4s  spill st %4s ld %4s  smem %s" % (name[:72], m.group(5), m.group(2), m.group(3), m.group(4), m.group(6) or 0))
