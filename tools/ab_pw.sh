#!/bin/bash
for lib in "$@"; do
  for w in piecewise3 piecewise4; do
  HGWARP_LIB=$PWD/homography.js_b200/$lib python bench.py --workload $w --steps 20 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib $w', round(d['value']), d['parity_gate'], round(d['fused']['pixel_kernel_ms_per_step'],4), round(d['roofline_frac_pixel_kernel'],3))"
  done
done
