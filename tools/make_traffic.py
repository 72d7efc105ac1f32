#!/usr/bin/env python
"""profiles/traffic.json from the ncu --set full captures of the headline kernels (what bench.py reports as roofline.traffic).
   python tools/make_traffic.py gpurun_out/<tag> <frames_per_launch> <name for profiles/>"""
import csv, io, json, os, subprocess, sys
src, frames, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
out = json.load(open(dst)) if os.path.exists(dst) else {}   # kernels without a new capture keep their entry (and its commit)
out["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full captures; bench.py scales them to its "
               "frames per launch; every entry names the commit its capture was taken at")
out["commit"] = commit
for rep, kernel in (("geo_projective", "warp_inverse_geo_kernel<projective>"), ("geo_affine", "warp_inverse_geo_kernel<affine>")):
    path = os.path.join(src, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, r = rows[0], rows[1], rows[2]
    def val(name):
        i = h.index(name)
        v = float(r[i].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u[i]]
    out[kernel] = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")), "frames_per_launch": frames,
                   "capture": f"{tag}_{rep}_ncu.txt", "commit": commit}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
