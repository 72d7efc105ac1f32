#!/bin/bash
# config 4: frames per launch chain of the two-lane stream.  Usage under gpurun: bash tools/ab_c4_chunk.sh
for cfg in "64 128" "96 192" "128 256"; do
  set -- $cfg
  HG_PW_CHUNK=$1 python bench.py --workload config4 --steps 5 --warmup 3 --c4-slots $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c4 chunk=$1 slots=$2', round(d['value']), d.get('checksum_gate'), 'whole', round(d['roofline_frac_whole_step'],4))"
done
