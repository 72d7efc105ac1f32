#!/bin/bash
# A/B of library builds on the fine-mesh piecewise workloads (span binning pass).  Usage under gpurun: bash tools/ab_span.sh <tag> libA.so ...
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q -k "piecewise or stream or mesh or index or fused or config" 2>&1 | tail -3 | tee $out/pytest.txt
for lib in "$@"; do
  for w in config4 piecewise4; do
    extra=""; [ $w = piecewise4 ] && extra="--pw-frames 64"
    HGWARP_LIB=$PWD/homography.js_b200/$lib python bench.py --workload $w --steps 5 --warmup 3 $extra 2>$out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read()); g = d.get('parity_gate', d.get('checksum_gate'))
print('$lib $w', round(d['value']), g, 'whole', round(d['roofline_frac_whole_step'],4), 'pixel', round(d.get('roofline_frac_pixel_kernel') or 0,4))" | tee -a $out/ab.txt
  done
done
