#!/bin/bash
# launch list (durations) of one bench workload.  Usage: bash tools/gpu_launches.sh <tag> <name> <bench args...>
tag=$1; name=$2; shift 2
out=gpurun_out/$tag
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${name}_launches.csv python bench.py "$@" --steps 1 --warmup 3 > $out/${name}_launches.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$out/${name}_launches.csv", errors="ignore")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
seq = [(r[kn].split("(")[0][-40:], float(r[mv].replace(",", "")) / 1000.0) for r in rows[hdr + 1:] if len(r) > mv]
# the last step = the launches after the last pw_stream_frames / pw_setup kernel group: print the tail
print("last 14 launches (us):")
for k, v in seq[-14:]:
    print(f"   {v:9.1f}  {k}")
PY
