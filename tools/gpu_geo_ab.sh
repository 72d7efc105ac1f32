#!/bin/bash
# A/B of the affine / projective pixel loop: register gathers vs asynchronous gathers (HG_GEO_ASYNC=1).  Usage under gpurun: bash tools/gpu_geo_ab.sh [tag]
tag=${1:-geoab}
out=gpurun_out/$tag
mkdir -p $out
HG_GEO_ASYNC=1 python -m pytest tests -m gpu -x -q -k "not piecewise and not stream and not jpeg and not png" 2>&1 | tail -4 | tee $out/pytest_async.txt
for mode in sync async; do
  for w in projective affine projective_generic affine_rot90; do
    if [ $mode = async ]; then export HG_GEO_ASYNC=1; else unset HG_GEO_ASYNC; fi
    python bench.py --workload $w --steps 20 --warmup 5 --no-secondary --cpu-budget 1 --e2e-frames 2 > $out/${w}_$mode.json 2> $out/${w}_$mode.err || tail -3 $out/${w}_$mode.err
    python - <<PY
import json
try:
    d = json.load(open("$out/${w}_$mode.json")); r = d["roofline"]
    ex = r.get("same_points_every_frame") or {}
    print(f"$mode $w value={d['value']:.0f} frac={r['frac']:.3f} kernel_ms={r['avg_kernel_ms']:.4f} exact_frac={ex.get('frac')} parity={d['parity_gate']}")
except Exception as e:
    print("$mode $w: no line", e)
PY
  done
done
