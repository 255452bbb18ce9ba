#!/bin/bash
mkdir -p gpurun_out
for d in 0 1; do
  echo "=== PS_CONV_DEBUG=$d"
  PS_CONV_DEBUG=$d timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_nets_gpu.py -q -m gpu 2>&1 | tail -12
done
