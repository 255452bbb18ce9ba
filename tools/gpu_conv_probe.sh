#!/bin/bash
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_nets_gpu.py -q -m gpu 2>&1 | tail -4
for d in 0 16; do PS_CONV_DEBUG=$d timeout 120 python tools/bench_conv.py 2>&1 | tail -1; done
for d in 0 16; do PS_CONV_DEBUG=$d timeout 120 python tools/bench_conv.py --heavy 2>&1 | tail -1; done
