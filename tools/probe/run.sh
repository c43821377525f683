./build/pattern_probe > gpurun_out/pattern_probe_r1b.txt 2>&1; grep "tiles" gpurun_out/pattern_probe_r1b.txt
python tools/microbench.py --config b --iters 40 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)['pieces']
print(' '.join('%s=%.1f'%(k.split('.',1)[1], d[k]['us_median']) for k in d if 'nnz' not in k and 'build' not in k))"
