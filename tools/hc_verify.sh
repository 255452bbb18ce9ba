mkdir -p gpurun_out/verify
python tools/repro_sampler.py --batch 128 --view 0 --save gpurun_out/verify/in128.npz 2>&1 | tail -1
echo "=== soak 128 zero-cache"; python tools/repro_sampler.py --load gpurun_out/verify/in128.npz --iters 300 --zero-cache 2>&1 | tail -3
rm -f gpurun_out/verify/in128.npz
echo "=== pytest"; python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== bench"; python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; tail -2 gpurun_out/bench_r02c.err
echo "=== trace"; python tools/trace_lmconv.py 2>&1 | sed -n 1,12p
