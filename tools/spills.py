#!/usr/bin/env python3
"""List the kernels of the last build that spill, with their register counts (from the ptxas -v logs)."""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "lte-gnu-radio-code_b200", "csrc", "build")
for f in sorted(glob.glob(os.path.join(sys.argv[1] if len(sys.argv) > 1 else root, "*.ptxas.log"))):
    name = None
    for line in open(f):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and int(m.group(2)) > 0:
            dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().replace("ofdm::", "")
            print(os.path.basename(f), m.group(2), "B spilled:", dn.split("(")[0])
