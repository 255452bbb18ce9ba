#!/bin/bash
# After the ring-phase fix: sampler-only soaks (tokens must be identical every iteration), pipeline soaks, GPU tests.
mkdir -p gpurun_out/verify
python tools/repro_sampler.py --batch 128 --view 0 --save gpurun_out/verify/in128.npz 2>&1 | tail -1
python tools/repro_sampler.py --batch 64 --view 4 --save gpurun_out/verify/in64.npz 2>&1 | tail -1
echo "=== soak 128"; python tools/repro_sampler.py --load gpurun_out/verify/in128.npz --iters 600 2>&1 | tail -4
echo "=== soak 128 zero-cache"; python tools/repro_sampler.py --load gpurun_out/verify/in128.npz --iters 400 --zero-cache 2>&1 | tail -4
echo "=== soak 64 view 4 zero-cache"; python tools/repro_sampler.py --load gpurun_out/verify/in64.npz --iters 600 --zero-cache 2>&1 | tail -4
rm -f gpurun_out/verify/in128.npz gpurun_out/verify/in64.npz
echo "=== pipeline soaks"; python tools/repro_fault.py --no-core --cases 128:0,64:0:1 --steps 60 --out gpurun_out/verify 2>&1 | tail -6
echo "=== pytest"; python -m pytest tests -x -q -m gpu 2>&1 | tail -15
