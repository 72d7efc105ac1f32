#!/bin/bash
# A/B of library builds on the piecewise workloads (HGWARP_LIB override).  Usage under gpurun: bash tools/ab_pw.sh libA.so libB.so ...
for lib in "$@"; do
  for w in piecewise3 piecewise4 config5; do
    extra=""; [ $w = config5 ] && extra="--c5-frames 2048"
    HGWARP_LIB=$PWD/homography.js_b200/$lib python bench.py --workload $w --steps 10 --warmup 3 $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); g = d.get('parity_gate', d.get('checksum_gate'))
print('$lib $w', round(d['value']), g, 'whole', round(d['roofline_frac_whole_step'],3), 'pixel', round(d['roofline_frac_pixel_kernel'],3))"
  done
done
