#!/bin/bash
python tools/time_small_losses.py 256 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:smooth_ -c 14 python tools/time_small_losses.py 256 2>&1 | grep -E "smooth_|gpu__time|dram__bytes|warps_active|issue_active|dram_throughput" | head -60
