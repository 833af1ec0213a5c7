#!/bin/bash
# 2+ GPUs: NCCL gradient-equality test, then the training bench (exposed all-reduce with the split encoder bucket)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-8}
N=${N:-2}
timeout 600 python -u -m pytest tests/test_gpu_dist.py tests/test_gpu_grad.py -m gpu --timeout 300 -x -q -p no:cacheprovider -rP > gpurun_out/tests_dist.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|skipped|^E  |^\[" gpurun_out/tests_dist.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --mode train --steps 10 --warmup 3 2>gpurun_out/dist_train.err | tail -n 1 > gpurun_out/dist_train_n$N.json
python - <<P
import json
d=json.loads(open('gpurun_out/dist_train_n$N.json').read())
print('N=$N train', d['value'], d['ms_per_step'], json.dumps(d['config'].get('grad_allreduce')))
P
tail -n 2 gpurun_out/dist_train.err
