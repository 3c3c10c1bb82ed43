for v in "$@"; do
  touch orb_slam_2_ros_b200/csrc/orb_search.cu
  make -C orb_slam_2_ros_b200/csrc -s -j8 EXTRA="$v" ../lib/liborb_b200.so > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "$v: $(python tools/bf_latency.py 2>&1 | tail -1)"
done
