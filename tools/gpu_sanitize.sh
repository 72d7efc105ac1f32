#!/bin/bash
# compute-sanitizer over the GPU tests that exercise round 2's kernels (small frames only; the 4K tests are deselected:
# the sanitizer runs kernels 10-50x slower).  Usage under gpurun: bash tools/gpu_sanitize.sh [tag]
tag=${1:-sanitizer}
out=gpurun_out/$tag
mkdir -p $out
SEL='(irregular or windows_unrelated or triangle_soup or stream or forward or lattice or batch or piecewise_batch or config5_video or folded) and not 4k and not full_size and not config4'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > $out/memcheck.txt 2>&1
echo "memcheck rc=$?" | tee -a $out/memcheck.txt
grep -E "passed|failed|ERROR SUMMARY" $out/memcheck.txt | tail -3
SELR='(irregular or windows_unrelated or stream_under_both or stream_windows) and band or stream_under_both or forward_geometric_bit_exact or stream_windows_and_pixels'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SELR" > $out/racecheck.txt 2>&1
echo "racecheck rc=$?" | tee -a $out/racecheck.txt
grep -E "passed|failed|RACECHECK SUMMARY" $out/racecheck.txt | tail -3
