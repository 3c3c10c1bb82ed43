#!/bin/bash
for v in "$@"; do
  touch orb_slam_2_ros_b200/csrc/orb_hamming.cu
  make -C orb_slam_2_ros_b200/csrc -s -j8 EXTRA="$v" ../lib/liborb_b200.so > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  python bench.py --no-cpu --e2e-callers 1 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());h=d['hamming'];print('$v', round(h['value']/1e9,1),'Gcmp/s', round(h['roofline']['peak'],1), round(h['roofline']['frac'],3), h['planted_top1_found'])"
done
touch orb_slam_2_ros_b200/csrc/orb_hamming.cu
