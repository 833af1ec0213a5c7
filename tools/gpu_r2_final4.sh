#!/bin/bash
# last call of round 2: full GPU suite + smoke on the final binary, then the N=1 default line (20 steps) and the few-stream records
bash tools/gpu_tests.sh
bash tools/gpu_r2_final3.sh
