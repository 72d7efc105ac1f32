#!/bin/bash
# Secondary pixel kernels (forward scatter / gather, bilinear extension): bench lines + an ncu launch list with DRAM bytes.
# Usage under gpurun: bash tools/gpu_secondary.sh [outdir]
out=${1:-gpurun_out/s3}
mkdir -p $out
for w in affine_forward projective_bilinear; do
  python bench.py --workload $w --steps 20 --warmup 3 > $out/$w.json 2> $out/$w.err || tail -5 $out/$w.err
  cat $out/$w.json
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:"forward|bilinear" -c 12 --csv --log-file $out/${w}_ncu.csv \
      python bench.py --workload $w --steps 1 --warmup 3 --frames 2 > $out/${w}_ncu.log 2>&1 || tail -5 $out/${w}_ncu.log
done
