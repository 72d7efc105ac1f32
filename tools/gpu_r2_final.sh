#!/bin/bash
# Round-end record on one GPU: the GPU suite, smoke(), the default bench line with the driver's arguments, the reference arm,
# one ncu --set full capture of the headline kernel (+ the tall affine instance) and the headline's launch list.
# Usage under gpurun: bash tools/gpu_r2_final.sh [tag]
tag=${1:-r02_fin}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/smoke.txt
python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err || tail -20 $out/bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/reference.json 2> $out/reference.err || tail -5 $out/reference.err
python - <<PY
import json
try:
    d = json.load(open("$out/bench.json"))
    r = d["roofline"]
    print(f"value={d['value']:.0f} Mpix/s frac={r['frac']:.4f} kernel_ms={r['avg_kernel_ms']:.4f} e2e={d['e2e']['value']:.0f} pipe={d['e2e']['pipe']['value']:.0f} clocks={d['clocks']}")
    for s in d.get("secondary", []):
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items() if k in ("name", "value", "roofline_frac_whole_step", "roofline_frac_pixel_kernel", "parity_gate", "checksum_gate", "error")})
    print("reference:", open("$out/reference.json").read()[:300])
except Exception as e:
    print("no bench line:", e)
PY
cap() {  # cap <name> <kernel regex> <skip> <bench args...>
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -f -o $out/$name \
      python bench.py "$@" --steps 1 --warmup 3 --no-secondary --cpu-budget 1 > $out/$name.log 2>&1 || tail -3 $out/$name.log
  [ -f $out/$name.ncu-rep ] && python tools/ncu_summary.py $out/$name.ncu-rep 12 > $out/${name}_ncu.txt 2>&1
}
cap geo_projective 'warp_inverse_geo_kernel' 4 --workload projective
cap geo_rot90 'warp_inverse_geo_affine_tall_kernel' 4 --workload affine_rot90
rm -f $out/geo_rot90.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_projective.csv \
    python bench.py --workload projective --steps 1 --warmup 3 --no-secondary --cpu-budget 1 > $out/launches_projective.log 2>&1
head -30 $out/geo_projective_ncu.txt
