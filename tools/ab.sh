#!/bin/bash
# A/B: run the short bench against several builds of the library (HGWARP_LIB override)
for lib in "$@"; do
  HGWARP_LIB=$PWD/homography.js_b200/$lib python bench.py $BENCH_ARGS --steps 60 --warmup 5 --no-cpu-baseline --e2e-frames 1 > gpurun_out/ab_$lib.json 2> gpurun_out/ab_$lib.err || tail -3 gpurun_out/ab_$lib.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$lib.json")); r = d["roofline"]
print(f"$lib value={d['value']:.0f} frac={r['frac']:.3f} kernel_ms={r['avg_kernel_ms']:.4f} parity={d['config']['parity_gate']}")
PY
done
