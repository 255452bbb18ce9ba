#!/bin/bash
# pipelined bench sweep: sampler partition x halo group
for cfg in "24 4" "24 8" "16 4" "16 8" "16 16" "24 16"; do
  set -- $cfg
  PS_TC_HALO_GROUP=$2 timeout 300 python bench.py --sampler-sms $1 --steps 20 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('S=$1 halo_group=$2: value %.0f  e2e %.0f  ms/step %.2f  serial %.2f ms  sampler(serial) %.2f ms' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['one_step_at_a_time']['ms_per_step'], d['rooflines']['lmconv_tc_kernel']['ms_per_step']))"
done
