#!/bin/bash
# the driver's multi-GPU launch at N GPUs (default 2), shortened secondary workloads.  Usage under gpurun --gpus N: bash tools/gpu_n2.sh <tag> [N] [bench args]
tag=${1:-n2}; n=${2:-2}; shift 2
out=gpurun_out/$tag
mkdir -p $out
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $n --steps 10 --warmup 3 "$@" > $out/bench_n$n.json 2> $out/bench_n$n.err || tail -30 $out/bench_n$n.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$out/bench_n$n.json") if l.startswith("{")][-1])
except Exception as e:
    print("no bench line:", e); raise SystemExit
r = d["roofline"]
print(f"N={d['n_gpus']} value={d['value']:.0f} frac={r['frac']:.3f} e2e={d['e2e']['value']:.0f} pipe={d['e2e']['pipe']['value']:.0f} link={d['e2e']['host_link']['bidir_GBps_all_gpus']:.1f} pipe_frac={d['e2e']['pipe_frac_of_host_link']:.2f}")
for s in d.get("secondary", []):
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items() if k in ("name", "value", "ms_per_step", "roofline_frac_whole_step", "roofline_frac_pixel_kernel", "parity_gate", "checksum_gate", "error", "nccl_broadcast_ms", "checksum_of_checksums")}, s.get("e2e", {}).get("value"))
PY
