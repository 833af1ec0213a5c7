#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 1200 python -u -m pytest tests -m gpu --timeout 300 -q -x -p no:cacheprovider > gpurun_out/tests_pdl.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tests_pdl.log
for f in 1 0; do echo "== CUM_PDL=$f"; CUM_PDL=$f timeout 300 python tools/pruned_probe.py 2>&1 | grep "B=1\|B=4"
CUM_PDL=$f timeout 300 python bench.py --mode stream --model e6 --streams-total 4096 --hops 1 --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stream 4096 h1', d['ms_per_step'])"
CUM_PDL=$f timeout 300 python bench.py --mode stream --model e6 --streams-total 512 --hops 1 --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stream 512 h1', d['ms_per_step'])"
CUM_PDL=$f timeout 300 python bench.py --mode stream --model e6 --streams-total 1 --hops 1 --steps 50 --graph 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stream 1 h1 graph', d['ms_per_step'])"
CUM_PDL=$f timeout 300 python bench.py --no-variants --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('offline', d['ms_per_step'], d['value'])"
done
