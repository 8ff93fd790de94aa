#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of a .cubin / .so (no GPU needed):
    python profiles/tools/sass_hist.py <file> <substring of the mangled kernel name> [top]"""
import collections
import re
import subprocess
import sys

path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur, hist, total = None, collections.Counter(), 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or key not in cur:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", line)
    if m:
        hist[m.group(1)] += 1
        total += 1
print("total", total)
for op, n in hist.most_common(top):
    print("%6d %s" % (n, op))
