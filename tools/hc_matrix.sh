for b in 2 32; do echo "== batch $b"; timeout 120 python tools/bench_lmconv.py --batch $b --reps 2 2>&1 | grep -E "FAULT|waiter|sampler_ms" | cut -c1-250 | head -20; done
echo "== old epilogue"; PS_TC_DEBUG=32768 timeout 120 python tools/bench_lmconv.py --reps 2 2>&1 | grep -E "FAULT|sampler_ms" | cut -c1-250
timeout 600 python -m pytest tests/test_lmconv_gpu.py -x -q -m gpu -s 2>&1 | tail -8
