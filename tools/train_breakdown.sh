timeout 300 python bench.py --mode train --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})"
