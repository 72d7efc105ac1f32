# forward loops: parity tests + the three forward bench lines, under a few winner-plane budgets (HG_FWD_PLANE_MB)
python -m pytest tests -m gpu -x -q -k "forward or flows or lattice" 2>&1 | tail -3
for mb in ${FWD_MBS:-48}; do
for w in affine_forward affine_forward_general piecewise_forward; do
  HG_FWD_PLANE_MB=$mb python bench.py --workload $w --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('plane_mb=$mb $w', round(d['value']), d['parity_gate'], 'whole', round(d['roofline_frac_whole_step'],3), 'ms', round(d['ms_per_step'],3))"
done
done
