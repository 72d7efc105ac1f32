#!/bin/bash
# A/B of the mask-based run-record builder (and the two-lane default for span-binned meshes) against the previous build.
# Usage under gpurun: bash tools/ab_records.sh [tag]
tag=${1:-r02_rec}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/pytest.txt
line() {
  python -c "
import json,sys
d=json.loads(sys.stdin.read()); g = d.get('parity_gate', d.get('checksum_gate'))
print('$1', round(d['value']), g, 'whole', round(d['roofline_frac_whole_step'],4), 'pixel', round(d.get('roofline_frac_pixel_kernel') or 0,4))" | tee -a $out/ab.txt
}
run() {
  label=$1; lib=$2; shift 2
  for w in piecewise3 piecewise4 config4 config5; do
    extra=""; [ $w = config5 ] && extra="--c5-frames 4096 --c5-slots 512"
    env HGWARP_LIB=$PWD/homography.js_b200/$lib "$@" python bench.py --workload $w --steps 8 --warmup 3 $extra 2>$out/err.txt | line "$label $w"
  done
}
run old libhgwarp_old.so HG_PW_CHUNK=64
run new_same_policy libhgwarp.so HG_PW_LANES=0 HG_PW_CHUNK=64
run new_default libhgwarp.so HG_X=1
