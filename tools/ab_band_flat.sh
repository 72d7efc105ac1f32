#!/bin/bash
# A/B of the band binning pass with flattened (segment, row) pairs against the previous build, plus launch-chain settings.
# Usage under gpurun: bash tools/ab_band_flat.sh [tag]
tag=${1:-r02_flat}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/pytest.txt
line() {
  python -c "
import json,sys
d=json.loads(sys.stdin.read()); g = d.get('parity_gate', d.get('checksum_gate'))
print('$1', round(d['value']), g, 'whole', round(d['roofline_frac_whole_step'],4), 'pixel', round(d.get('roofline_frac_pixel_kernel') or 0,4))" | tee -a $out/ab.txt
}
for lib in libhgwarp_old.so libhgwarp.so; do
  for w in piecewise3 config5; do
    extra=""; [ $w = config5 ] && extra="--c5-frames 4096"
    HGWARP_LIB=$PWD/homography.js_b200/$lib python bench.py --workload $w --steps 10 --warmup 3 $extra 2>$out/err.txt | line "$lib $w"
  done
done
HG_PW_CHUNK=128 python bench.py --workload piecewise3 --steps 10 --warmup 3 --pw-frames 128 2>$out/err.txt | line "new pw3 128 frames chunk 128"
for sl in 512 2048; do
  python bench.py --workload config5 --steps 5 --warmup 3 --c5-slots $sl 2>$out/err.txt | line "new c5 12500 frames slots=$sl"
done
python bench.py --workload config4 --steps 5 --warmup 3 2>$out/err.txt | line "new c4 serial"
for ch in 32 64; do
  HG_PW_LANES=1 HG_PW_CHUNK=$ch python bench.py --workload config4 --steps 5 --warmup 3 2>$out/err.txt | line "new c4 lanes chunk=$ch"
done
HG_PW_LANES=1 HG_PW_CHUNK=64 python bench.py --workload config4 --steps 5 --warmup 3 --c4-slots 256 2>$out/err.txt | line "new c4 lanes chunk=64 slots=256"
