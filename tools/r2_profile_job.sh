#!/bin/bash
# Round-2 profiling job (run under gpurun): ncu launch lists with DRAM/L2 metrics of the three phases, one
# --set full capture of the dominant kernel of the index build, compute-sanitizer memcheck + racecheck.
set -x
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct
for ph in index scans fm; do
  python tools/ncu_step.py --profile $ph > gpurun_out/r2b_plain_$ph.log 2>&1 || exit 1
  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2b_ncu_$ph.csv --metrics $M \
      python tools/ncu_step.py --profile $ph > gpurun_out/r2b_ncu_$ph.log 2>&1
done
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bucket_sort_kernel -c 1 \
    -o gpurun_out/prof_r2_bucket_sort python tools/ncu_step.py --profile index > gpurun_out/r2b_ncu_full.log 2>&1
ncu -i gpurun_out/prof_r2_bucket_sort.ncu-rep --page raw --csv > gpurun_out/prof_r2_bucket_sort_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2_bucket_sort.ncu-rep --page details > gpurun_out/prof_r2_bucket_sort_details.txt 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_step.py > gpurun_out/r2b_sanitize_memcheck.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_step.py > gpurun_out/r2b_sanitize_racecheck.log 2>&1
tail -4 gpurun_out/r2b_sanitize_memcheck.log gpurun_out/r2b_sanitize_racecheck.log
