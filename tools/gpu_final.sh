#!/bin/bash
# Round-end check on the GPU box: the whole GPU suite, smoke(), the bilinear line and the headline bench line.
# Usage under gpurun: bash tools/gpu_final.sh [outdir]
out=${1:-gpurun_out/final}
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --workload projective_bilinear --steps 20 --warmup 3 > $out/projective_bilinear.json 2> $out/projective_bilinear.err || tail -5 $out/projective_bilinear.err
cat $out/projective_bilinear.json
python bench.py --steps 100 --warmup 5 --cpu-budget 4 > $out/bench.json 2> $out/bench.err || tail -5 $out/bench.err
cat $out/bench.json
