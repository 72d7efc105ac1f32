#!/bin/bash
# A/B of library builds on the affine / projective pixel loop (HGWARP_LIB override) + the GPU suite on the default build.
# Usage under gpurun: [SKIP_TESTS=1] [WORKLOADS="projective ..."] bash tools/ab_geo_lib.sh <tag> libA.so libB.so ...
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
[ -z "$SKIP_TESTS" ] && python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest.txt
for lib in "$@"; do
  for w in ${WORKLOADS:-projective affine projective_generic affine_rot90}; do
    HGWARP_LIB=$PWD/homography.js_b200/$lib python bench.py --workload $w --steps 20 --warmup 5 --no-secondary --cpu-budget 1 --e2e-frames 2 > $out/${w}_$lib.json 2> $out/${w}_$lib.err || tail -3 $out/${w}_$lib.err
    python - <<PY | tee -a $out/ab.txt
import json
try:
    d = json.load(open("$out/${w}_$lib.json")); r = d["roofline"]
    ex = r.get("same_points_every_frame") or {}
    print(f"$lib $w value={d['value']:.0f} frac={r['frac']:.4f} kernel_ms={r['avg_kernel_ms']:.4f} exact_frac={ex.get('frac')} parity={d['parity_gate']}")
except Exception as e:
    print("$lib $w: no line", e)
PY
  done
done
