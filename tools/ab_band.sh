#!/bin/bash
# A/B of the piecewise binning passes: span + run kernels over global bins (HG_PW_BINNING=1) against the one-pass band
# kernel through shared memory (2).  Usage under gpurun: bash tools/ab_band.sh [tag]
tag=${1:-band}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q -k "piecewise or stream or fused or flows or config or pipe or mesh or index" 2>&1 | tail -8 | tee $out/pytest.txt
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
run span HG_PW_BINNING=1
run band HG_PW_BINNING=2
