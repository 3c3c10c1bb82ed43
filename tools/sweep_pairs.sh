for v in "$@"; do
  touch orb_slam_2_ros_b200/csrc/orb_search.cu
  make -C orb_slam_2_ros_b200/csrc -s -j8 EXTRA="$v" ../lib/liborb_b200.so > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  python bench.py --no-hamming --no-cpu --e2e-callers 1 --steps 10 --warmup 3 > /tmp/sw.json 2>/dev/null
  python -c "
import json;d=json.load(open('/tmp/sw.json'));p=d['pairs']['config2'];q=d['pairs']['config3'];print('$v',round(p['value']),round(p['ms_per_step'],3),round(p['matching_only_ms_per_step'],3),round(q['value']),round(q['stereo_matching_only_ms_per_step'],3))"
done
