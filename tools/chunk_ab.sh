for ch in 16 32 64 1024; do
  for w in piecewise3 piecewise4; do
    HG_PW_CHUNK=$ch python bench.py --workload $w --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chunk $ch $w', round(d['value']), 'whole', round(d['roofline_frac_whole_step'],3), 'pixel', round(d['roofline_frac_pixel_kernel'],3), 'ms', round(d['ms_per_step'],3))"
  done
done
