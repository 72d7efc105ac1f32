#!/usr/bin/env python
"""Per-kernel SASS listing / opcode histogram of libhgwarp.so:  python tools/sass.py <regex> [--list]"""
import collections
import re
import subprocess
import sys

lib = "homography.js_b200/libhgwarp.so"
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else re.compile(".")
out = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(line)
for name, lines in funcs.items():
    if not pat.search(name):
        continue
    print("==", name, len(lines), "instructions")
    if "--list" in sys.argv:
        for l in lines:
            print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l))
    else:
        h = collections.Counter()
        for l in lines:
            ins = re.sub(r"^\s+/\*[0-9a-f]{4}\*/\s+", "", l)
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            h[ins.split()[0].rstrip(";").split(".")[0]] += 1
        print("  ".join(f"{k}:{v}" for k, v in h.most_common()))
