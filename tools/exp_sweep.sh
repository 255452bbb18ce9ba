#!/bin/bash
# pipelined bench sweep: batch x sampler partition x depth
for cfg in "128 24 2" "128 32 2" "128 40 2" "96 32 2" "64 24 3" "32 24 2"; do
  set -- $cfg
  timeout 300 python bench.py --batch $1 --sampler-sms $2 --in-flight $3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('batch $1 S=$2 depth $3: value %.0f  e2e %.0f  ms/step %.2f  serial %.0f  conv %.2f ms  sampler %.2f ms' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['one_step_at_a_time']['value'], d['rooflines']['conv_igemm_kernel']['ms_per_step'], d['rooflines']['lmconv_tc_kernel']['ms_per_step']))"
done
