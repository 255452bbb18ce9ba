mkdir -p gpurun_out/verify
python tools/repro_sampler.py --batch 128 --view 0 --save gpurun_out/verify/in128.npz 2>&1 | tail -1
echo "=== soak 128 zero-cache"; python tools/repro_sampler.py --load gpurun_out/verify/in128.npz --iters 300 --zero-cache 2>&1 | tail -3
rm -f gpurun_out/verify/in128.npz
echo "=== trace"; python tools/trace_lmconv.py 2>&1 | head -40
echo "=== pytest"; python -m pytest tests -x -q -m gpu 2>&1 | tail -5
