#!/bin/bash
# One `ncu --set full` capture per pixel kernel + a launch list of the headline bench, summarised to text.
# Usage under gpurun (one GPU; ~3 min of box time):  bash tools/gpu_profile_all.sh [tag]     -> gpurun_out/<tag>/
# Numbers printed by runs under ncu are never bench values; the bench lines come from tools/gpu_final.sh.
tag=${1:-prof}
out=gpurun_out/$tag
mkdir -p $out
cap() {  # cap <name> <kernel regex> <skip> <bench args...>
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -f -o $out/$name \
      python bench.py "$@" --steps 1 --warmup 3 --no-cpu-baseline > $out/$name.log 2>&1 || tail -3 $out/$name.log
  [ -f $out/$name.ncu-rep ] && python tools/ncu_summary.py $out/$name.ncu-rep 12 > $out/${name}_ncu.txt 2>&1
  # gpurun merges at most 64 MiB back: keep the raw report of the kernels most likely to be read line by line
  case $name in geo_projective|pw4_pixel|pw4_span|bilinear) ;; *) rm -f $out/$name.ncu-rep ;; esac
}
cap geo_projective   'warp_inverse_geo_kernel'    3 --workload projective
cap geo_affine       'warp_inverse_geo_kernel'    3 --workload affine
cap geo_rot90        'warp_inverse_geo_kernel'    3 --workload affine_rot90
cap pw3_pixel        'pw_warp_fused_kernel'       3 --workload piecewise3
cap pw3_span         'pw_span_bin_kernel'         3 --workload piecewise3
cap pw3_runs         'pw_bin_runs_kernel'         3 --workload piecewise3
cap pw4_pixel        'pw_warp_fused_kernel'       3 --workload piecewise4
cap pw4_span         'pw_span_bin_kernel'         3 --workload piecewise4
cap pw4_runs         'pw_bin_runs_kernel'         3 --workload piecewise4
cap bilinear         'bilinear2_kernel'           3 --workload projective_bilinear
cap fwd_scatter      'forward_scatter_kernel'   200 --workload affine_forward
cap fwd_gather       'forward_gather_kernel'    200 --workload affine_forward
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --cpu-budget 1 > $out/launches_bench.log 2>&1
ls -la $out | head -40
