#!/bin/bash
# round-2 GPU loop: parity tests + the full default bench line (headline + secondary list).  Usage under gpurun: bash tools/gpu_r2.sh [tag] [bench args]
tag=${1:-r2}; shift
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest.txt
python bench.py "$@" > $out/bench.json 2> $out/bench.err || tail -20 $out/bench.err
python - <<PY
import json
try:
    d = json.load(open("$out/bench.json"))
except Exception as e:
    print("no bench line:", e); raise SystemExit
r = d["roofline"]
print(f"value={d['value']:.0f} Mpix/s frac={r['frac']:.3f} kernel_ms={r['avg_kernel_ms']:.4f} e2e={d['e2e']['value']:.0f} pipe={d['e2e']['pipe']['value']:.0f} link={d['e2e']['host_link']} clocks={d['clocks']}")
for s in d.get("secondary", []):
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items() if k in ("name", "value", "ms_per_step", "roofline_frac_whole_step", "roofline_frac_pixel_kernel", "parity_gate", "checksum_gate", "error", "nccl_broadcast_ms", "frames_fused_vs_general")}, s.get("e2e", {}).get("value"))
PY
