#!/bin/bash
# Captures the ncu evidence of one round on the GPU box (run through gpurun from the repo root):
#   bash profiles/capture.sh r1
# then, back in the build container:  python profiles/summarize.py r1
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
# plain run first: the profiled command must exit 0 without ncu, and tells the launches per step
python profiles/run_step.py cfg3 2 > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
N=$(tail -n 1 $OUT/plain_$TAG.log | awk '{print $4}')
echo "launches per step: $N"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum"
M="$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
M="$M,launch__registers_per_thread"
ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file $OUT/launches_$TAG.csv \
    python profiles/run_step.py cfg3 2 > $OUT/ncu_launches_$TAG.log 2>&1
ncu --metrics $M --clock-control none -s $N -c $N --csv --log-file $OUT/metrics_$TAG.csv \
    python profiles/run_step.py cfg3 2 > $OUT/ncu_metrics_$TAG.log 2>&1
# level-1 launches (64 pairs) of the top kernels, full set with source correlation
ncu --set full --clock-control none --import-source on \
    -k regex:"k_mc_march|k_subpel_tma|k_subpel_strip|k_upsample_chain" -c 8 -o $OUT/top_$TAG -f \
    python profiles/run_step.py cfg3 1 > $OUT/ncu_top_$TAG.log 2>&1
# the dominant kernel (its launches come after the eight above)
ncu --set full --clock-control none --import-source on -k regex:k_mc_march -c 2 -o $OUT/march_$TAG -f \
    python profiles/run_step.py cfg3 1 > $OUT/ncu_march_$TAG.log 2>&1
ls -la $OUT/*_$TAG.*
