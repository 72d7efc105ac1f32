#!/bin/bash
# last record of the round on one GPU: the GPU suite, smoke(), the default bench line with the driver's arguments.
# Usage under gpurun: bash tools/gpu_r2_last.sh [tag]
tag=${1:-r02_last}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/smoke.txt
python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err || tail -20 $out/bench.err
python - <<PY
import json
d = json.load(open("$out/bench.json"))
r = d["roofline"]
print(f"value={d['value']:.0f} Mpix/s frac={r['frac']:.4f} kernel_ms={r['avg_kernel_ms']:.4f} e2e={d['e2e']['value']:.0f} pipe={d['e2e']['pipe']['value']:.0f} clocks={d['clocks']}")
for s in d.get("secondary", []):
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items() if k in ("name", "value", "roofline_frac_whole_step", "roofline_frac_pixel_kernel", "parity_gate", "checksum_gate", "error")})
PY
