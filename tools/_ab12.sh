for g in 4 8 16 32; do
  python bench.py --steps 256 --warmup 16 --no-cpu-baseline --graph-steps $g 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.load(sys.stdin); print('G=$g bench value=%.0f ms=%.4f' % (d['value'], d['ms_per_step']))"
done
