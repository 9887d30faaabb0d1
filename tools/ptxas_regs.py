"""Kernel name -> registers / spills from a ptxas -v log (sift_b200/csrc/_build/*.ptxas.log)."""
import re
import subprocess
import sys

t = open(sys.argv[1]).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores[^\n]*\n[^\n]*Used (\d+) registers", t):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name).replace("void siftgpu::", "")
    print(f"{name:48s} regs {m.group(4):>3s}  stack {m.group(2)}  spill {m.group(3)}")
