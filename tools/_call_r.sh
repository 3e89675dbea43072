#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r_launches_int8_bert.csv python tools/prof_int8_step.py > gpurun_out/r_int8_bert.log 2>&1
python tools/summarize_launches.py gpurun_out/r_launches_int8_bert.csv > gpurun_out/r_launches_int8_bert.summary.txt 2>&1; head -24 gpurun_out/r_launches_int8_bert.summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r_launches_int8_opt.csv python tools/prof_int8_step.py --opt > gpurun_out/r_int8_opt.log 2>&1
python tools/summarize_launches.py gpurun_out/r_launches_int8_opt.csv > gpurun_out/r_launches_int8_opt.summary.txt 2>&1; head -24 gpurun_out/r_launches_int8_opt.summary.txt
