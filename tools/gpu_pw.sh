#!/bin/bash
# piecewise loop: parity tests + the piecewise bench lines.  Usage under gpurun: bash tools/gpu_pw.sh [tag] [all]
tag=${1:-pw}
out=gpurun_out/$tag
mkdir -p $out
if [ "$2" = "all" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest.txt
else
  python -m pytest tests -m gpu -x -q -k "piecewise or stream or fused or flows or config or pipe or mesh or index" 2>&1 | tail -15 | tee $out/pytest.txt
fi
for w in piecewise3 piecewise4 config5; do
  extra=""; [ $w = config5 ] && extra="--c5-frames 2048"
  python bench.py --workload $w --steps 10 --warmup 3 $extra > $out/$w.json 2> $out/$w.err || tail -5 $out/$w.err
  python - <<PY
import json
try:
    d = json.load(open("$out/$w.json"))
    print("$w", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ("value", "ms_per_step", "roofline_frac_whole_step", "roofline_frac_pixel_kernel", "pixel_kernel_ms_per_step", "parity_gate", "checksum_gate", "error")})
except Exception as e:
    print("$w: no line", e)
PY
done
