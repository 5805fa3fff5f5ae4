#!/bin/bash
python tools/profile_step.py 32 3 vg > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:warp_photo_stream -s 1 -c 1 -f -o gpurun_out/prof_r2_final python tools/profile_step.py 32 3 vg > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log
python tools/ncu_phase_summary.py gpurun_out/prof_r2_final.ncu-rep > gpurun_out/prof_r2_final_summary.txt 2>&1
rm -f gpurun_out/prof_r2_final.ncu-rep
head -50 gpurun_out/prof_r2_final_summary.txt
