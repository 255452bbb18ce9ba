#!/bin/bash
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_nets_gpu.py -q -m gpu 2>&1 | tail -3
for d in 0 8 16 28; do echo "PS_CONV_DEBUG=$d"; PS_CONV_DEBUG=$d timeout 300 python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['rooflines']['conv_igemm_kernel']; print('  conv ms/step %.3f  step %.2f'%(r['ms_per_step'], d['ms_per_step']))"; done
