"""Summarise ptxas -v output (registers / spills per kernel) from csrc/ptxas.log."""
import re, subprocess, sys, os
log = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ptxas.log")).read()
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers")
for name, stack, ss, sl, regs in pat.findall(log):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(vrt::MarchParams\)|void vrt::", "", dem)
    print(f"{regs:>4} regs  stack {stack:>3}  spill {ss}/{sl}  {dem}")
