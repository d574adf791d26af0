for c in 1 2 3 4 6 8 12; do
  SPG_FAST_CTAS=$c python bench.py --sizes 4,5,8 --blankets 50000 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ctas', $c, [(p['n'], round(p['ms'],3)) for p in d['per_size']])"
done
