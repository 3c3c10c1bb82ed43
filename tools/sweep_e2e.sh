#!/bin/bash
# like sweep.sh, prints value / e2e only (single caller)
f=$1; shift
for v in "$@"; do
  touch orb_slam_2_ros_b200/csrc/$f
  make -C orb_slam_2_ros_b200/csrc -s -j8 EXTRA="$v" ../lib/liborb_b200.so > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  python bench.py --no-hamming --no-cpu --e2e-callers 1 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v', round(d['value']),round(d['e2e']['value']))"
done
touch orb_slam_2_ros_b200/csrc/$f
