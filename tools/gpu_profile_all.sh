#!/bin/bash
# One `ncu --set full` capture per kernel + launch lists of the bench workloads, summarised to text.
# Usage under gpurun (one GPU; ~5 min of box time):  bash tools/gpu_profile_all.sh [tag]     -> gpurun_out/<tag>/
# Numbers printed by runs under ncu are never bench values; the bench lines come from tools/gpu_r2.sh.
tag=${1:-prof}
out=gpurun_out/$tag
mkdir -p $out
cap() {  # cap <name> <kernel regex> <skip> <bench args...>
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -f -o $out/$name \
      python bench.py "$@" --steps 1 --warmup 3 --no-secondary --cpu-budget 1 > $out/$name.log 2>&1 || tail -3 $out/$name.log
  [ -f $out/$name.ncu-rep ] && python tools/ncu_summary.py $out/$name.ncu-rep 12 > $out/${name}_ncu.txt 2>&1
  # gpurun merges at most 64 MiB back: keep the raw report of the kernels most likely to be read line by line
  case $name in geo_projective|geo_affine|pw4_pixel|pw3_pixel) ;; *) rm -f $out/$name.ncu-rep ;; esac
}
# the headline batch launches per step: solve_kernel + warp_inverse_geo_kernel; parity gate + 3 warm-up steps come first
cap geo_projective   'warp_inverse_geo_kernel'    4 --workload projective
cap geo_affine       'warp_inverse_geo_kernel'    4 --workload affine
cap geo_rot90        'warp_inverse_geo_kernel'    4 --workload affine_rot90
cap pw3_pixel        'pw_warp_fused_kernel'       3 --workload piecewise3 --pw-frames 16
cap pw3_band         'pw_band_bins_kernel'        3 --workload piecewise3 --pw-frames 16
cap pw4_pixel        'pw_warp_fused_kernel'       3 --workload piecewise4 --pw-frames 16
cap pw4_span         'pw_span_bin_kernel'         3 --workload piecewise4 --pw-frames 16
cap pw4_runs         'pw_bin_runs_kernel'         3 --workload piecewise4 --pw-frames 16
cap c5_pixel         'pw_warp_fused_kernel'       8 --workload config5 --c5-frames 1024
cap c5_band          'pw_band_bins_kernel'        8 --workload config5 --c5-frames 1024
cap bilinear         'bilinear2_kernel'           4 --workload projective_bilinear
cap fwd_lattice      'forward_lattice_kernel'     4 --workload affine_forward
cap fwd_scatter      'forward_scatter_kernel'    20 --workload affine_forward_general
cap fwd_gather       'forward_gather_kernel'     20 --workload affine_forward_general
for w in projective piecewise3 piecewise4 config5 affine_forward_general piecewise_forward; do
  extra=""; [ $w = config5 ] && extra="--c5-frames 1024"; [ $w = piecewise3 -o $w = piecewise4 ] && extra="--pw-frames 16"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 3 --no-secondary --cpu-budget 1 $extra > $out/launches_$w.log 2>&1
done
ls -la $out | head -60
