#!/bin/bash
# sweep.sh file "-DA=1" "-DA=2" ... : rebuild (file = a .cu, or orb_internal.cuh for everything) with each flag set on
# the GPU box and print the stage timers of a short bench run
f=$1; shift
for v in "$@"; do
  touch orb_slam_2_ros_b200/csrc/$f
  make -C orb_slam_2_ros_b200/csrc -s -j8 EXTRA="$v" ../lib/liborb_b200.so > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  python bench.py --no-hamming --no-cpu --steps 10 --warmup 3 > /tmp/sw.json 2> /tmp/sw.err
  python -c "
import json;d=json.load(open('/tmp/sw.json'));print('$v',round(d['value']),round(d['e2e']['value']),{k:round(v['ms'],3) for k,v in d['roofline']['stages'].items()})"
done
touch orb_slam_2_ros_b200/csrc/$f
