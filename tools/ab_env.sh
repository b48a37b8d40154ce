#!/bin/bash
# A/B of one environment setting on the per-op profile: usage  tools/ab_env.sh NAME=VALUE [pattern]
for v in "A=0" "$1" "A=0" "$1"; do
  echo "== $v"; env $v timeout 300 python tools/profile_ops.py 32 0.3 2>&1 | grep -E "^batch|${2:-conv3x3   n=  1 }"
done
