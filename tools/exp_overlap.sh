#!/bin/bash
for m in "split2 8" "split2 16" "split2 24" "split2 32" "split2 40" "green 96"; do
  timeout 200 python tools/exp_overlap.py $m 2>&1 | grep -v Warning | tail -2
done
