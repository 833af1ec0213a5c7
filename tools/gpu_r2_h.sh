#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 OMP_NUM_THREADS=${OMP_NUM_THREADS:-16}
timeout 900 python -u -m pytest tests/test_gpu_model.py -m gpu -rP --timeout 300 -q -p no:cacheprovider -k "benchmark_clip or vs_exact_fp32 or golden" > gpurun_out/tests_model.log 2>&1; echo "pytest model rc=$?"
grep -E "^\[E8|passed|failed|^E  " gpurun_out/tests_model.log | tail -12
for f in 1 0; do CUM_FUSED_ENDS=$f timeout 300 python - <<PY
import torch, sys, json
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import bench
from cleanumamba_b200.network import Net
torch.manual_seed(0)
net = Net("CleanUMamba", dict(bench.CONFIGS["e8"], math_mode="f16x3")).cuda().eval()
ref = Net("CleanUMamba", dict(bench.CONFIGS["e8"], math_mode="fp32")).cuda().eval()
ref.load_state_dict(net.state_dict())
x = bench.synth_noisy(8, 10.0, 3).cuda()
with torch.no_grad():
    y = net(x.clone()); r = ref(x.clone())
print("fused_ends=$f  8 x 10 s f16x3 vs fp32 max-abs", (y-r).abs().max().item())
PY
done
