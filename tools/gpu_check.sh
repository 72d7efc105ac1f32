#!/bin/bash
# quick GPU loop: parity tests + a short bench line (key numbers only).  Usage under gpurun: bash tools/gpu_check.sh [tag]
tag=${1:-check}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 --cpu-budget 1 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
r = d["roofline"]
print(f"value={d['value']:.0f} Mpix/s  frac={r['frac']:.3f}  kernel_ms={r['avg_kernel_ms']:.4f}  e2e={d['e2e']['value']:.0f}  clocks={d['clocks']}")
PY
