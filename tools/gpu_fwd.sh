python -m pytest tests -m gpu -x -q -k "forward or flows or lattice" 2>&1 | tail -3
for w in affine_forward affine_forward_general piecewise_forward; do
  python bench.py --workload $w --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w', round(d['value']), d['parity_gate'], 'whole', round(d['roofline_frac_whole_step'],3), 'ms', round(d['ms_per_step'],3))"
done
