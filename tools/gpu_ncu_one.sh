#!/bin/bash
# one ncu --set full capture of one kernel of one bench workload, summarised.  Usage: bash tools/gpu_ncu_one.sh <tag> <name> <kernel regex> <skip> <bench args...>
tag=$1; name=$2; rx=$3; skip=$4; shift 4
out=gpurun_out/$tag
mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -f -o $out/$name \
    python bench.py "$@" --steps 1 --warmup 3 --no-cpu-baseline > $out/$name.log 2>&1 || tail -3 $out/$name.log
[ -f $out/$name.ncu-rep ] && python tools/ncu_summary.py $out/$name.ncu-rep 25 > $out/${name}_ncu.txt 2>&1
cat $out/${name}_ncu.txt | head -60
