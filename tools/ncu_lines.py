#!/usr/bin/env python
"""Executed warp instructions and stall samples per SOURCE LINE of a kernel in an .ncu-rep (needs -lineinfo + --import-source on).
   python tools/ncu_lines.py report.ncu-rep [n_top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, out = None, []
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0].isdigit() and r[2] == "-":
        try: out.append((int(r[7]), int(r[4]), cur, int(r[0]), r[1].strip()[:110], r[10]))
        except ValueError: pass
tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print(f"total warp instr {tot}, stall samples {ts}")
for o in sorted(out, reverse=True)[:ntop]:
    print(f"{100*o[0]/tot:5.1f}% instr {100*o[1]/ts:5.1f}% stall  {o[2]}:{o[3]:<4d} thr/warp={o[5]:>3s}  {o[4]}")
