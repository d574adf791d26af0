# A/B of the phase-aligned multi-warp CTAs of the sub-warp-group fast kernels (SPG_FAST_WPC=1: one warp per CTA as before)
for w in 1 0; do
  if [ $w = 1 ]; then export SPG_FAST_WPC=1; else unset SPG_FAST_WPC; fi
  python bench.py --sizes ${SIZES:-3,4,5,6,8} --blankets 100000 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('wpc', '$w' if '$w'=='1' else 'auto', [(p['n'], round(p['ms'],3)) for p in d['per_size']], 'ok', d.get('blankets_ok_device'))"
done
