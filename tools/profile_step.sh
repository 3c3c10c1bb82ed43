#!/bin/bash
# profile_step.sh <tag>: on the GPU box, the evidence files of ONE bench step (512 frames 640x480) of the current build:
#   gpurun_out/<tag>_launches.csv / .txt   ncu launch list (durations, instructions, pipe utilisation, DRAM bytes), no replay sections
#   gpurun_out/<tag>_full.ncu-rep          ncu --set full --import-source on of the same 12 launches
#   gpurun_out/<tag>_ncu_full.csv / .txt, <tag>_ncu_pipes.txt, <tag>_pipes.json, <tag>_traffic.json   summaries of it
#   gpurun_out/<tag>_fast_regions.txt      executed-instruction profile of fast_strip_kernel per 40-instruction SASS region
# The bench numbers themselves are NEVER taken from a run under ncu.
tag=${1:-prof}
K='regex:pyr_|fast_strip|quadtree|blur_kernel|orient_describe'
B="python bench.py --no-hamming --no-cpu --no-pairs --e2e-callers 1 --steps 2 --warmup 3"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k "$K" -s 36 -c 12 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
python tools/launch_list_summary.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_launches.txt "one bench step, 512 frames 640x480 (ncu --metrics)"
ncu --set full --import-source on --clock-control none -k "$K" -s 36 -c 12 -f -o gpurun_out/${tag}_full $B > gpurun_out/${tag}_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_full.ncu-rep gpurun_out/${tag}_ncu_full.csv > /dev/null
python tools/ncu_table.py gpurun_out/${tag}_ncu_full.csv > gpurun_out/${tag}_ncu_full.txt 2>&1
python tools/ncu_pipes.py gpurun_out/${tag}_full.ncu-rep > gpurun_out/${tag}_ncu_pipes.txt
python tools/ncu_pipes_json.py gpurun_out/${tag}_full.ncu-rep gpurun_out/${tag}_pipes.json > /dev/null
python tools/ncu_traffic.py gpurun_out/${tag}_ncu_full.csv 512 gpurun_out/${tag}_traffic.json > /dev/null
python tools/ncu_sass_regions.py gpurun_out/${tag}_full.ncu-rep fast_strip 40 > gpurun_out/${tag}_fast_regions.txt
cat gpurun_out/${tag}_launches.txt
