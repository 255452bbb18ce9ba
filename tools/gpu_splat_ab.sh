#!/bin/bash
# A/B of the splat tile kernel encodings: parity suite + map-mode timing per PS_SPLAT_VARIANT given as arguments.
mkdir -p gpurun_out
for v in "$@"; do
  PS_SPLAT_VARIANT=$v timeout 200 python -m pytest tests/test_splat_gpu.py -q -m gpu -x > gpurun_out/pytest_splat_v$v.log 2>&1
  echo "V=$v pytest rc=$? $(tail -1 gpurun_out/pytest_splat_v$v.log)"
  PS_SPLAT_VARIANT=$v timeout 100 python tools/bench_splat.py 2>&1 | tail -1
done
