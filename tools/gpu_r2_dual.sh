#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python tools/dual_stream_probe.py 4096 1 2>&1 | tail -8
timeout 600 python tools/dual_stream_probe.py 4096 2 2>&1 | tail -5
