#!/bin/bash
# A/B of the two-lane piecewise pipeline (binning passes of chunk k+1 beside the pixel kernel of chunk k) against the one-lane
# order.  Usage under gpurun: bash tools/ab_lanes.sh [tag]
tag=${1:-lanes}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q -k "piecewise or stream or fused or flows or config or pipe or mesh or index" 2>&1 | tail -5 | tee $out/pytest.txt
run() {
  label=$1; shift
  for w in piecewise3 piecewise4 config4 config5; do
    extra=""; [ $w = config5 ] && extra="--c5-frames 2048"
    env "$@" python bench.py --workload $w --steps 10 --warmup 3 $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); g = d.get('parity_gate', d.get('checksum_gate'))
print('$label $w', round(d['value']), g, 'whole', round(d['roofline_frac_whole_step'],3), 'pixel', round(d.get('roofline_frac_pixel_kernel') or 0,3))" | tee -a $out/ab.txt
  done
}
run serial64 HG_PW_SERIAL=1
run lanes16 HG_X=1
run lanes8 HG_PW_CHUNK=8
run lanes32 HG_PW_CHUNK=32
