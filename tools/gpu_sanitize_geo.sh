#!/bin/bash
# compute-sanitizer memcheck over the GPU tests of the affine / projective loop (incl. the tall affine instance) and of the band
# binning pass, small frames only.  Usage under gpurun: bash tools/gpu_sanitize_geo.sh [tag]
tag=${1:-sanitizer_geo}
out=gpurun_out/$tag
mkdir -p $out
SEL='(projective or boundar or horizon or fast_body or inverse_points or pipelined or affine or rotat or quarter_turn or irregular or windows_unrelated) and not 4k and not full_size and not config4 and not bilinear'
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > $out/memcheck.txt 2>&1
echo "memcheck rc=$?" | tee -a $out/memcheck.txt
grep -E "passed|failed|ERROR SUMMARY" $out/memcheck.txt | tail -3
