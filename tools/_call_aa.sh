#!/bin/bash
# GPU call aa: launch lists of the three eval steps at HEAD; ncu --set full of the persistent attention kernel (T=2048, B=32),
# the one-tile kernel (S row in registers) at the C-BERT shape and the erf GELU kernel
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/aa_launches_bert.csv python bench.py --steps 2 --warmup 3 --no-graph --no-sweep --no-cpu --profile-one-step > gpurun_out/aa_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/aa_launches_bert.csv > gpurun_out/aa_launches_bert.summary.txt 2>&1; head -16 gpurun_out/aa_launches_bert.summary.txt | cut -c1-160
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/aa_launches_int8_bert.csv python tools/prof_int8_step.py > gpurun_out/aa_int8_bert.log 2>&1
python tools/summarize_launches.py gpurun_out/aa_launches_int8_bert.csv > gpurun_out/aa_launches_int8_bert.summary.txt 2>&1; head -22 gpurun_out/aa_launches_int8_bert.summary.txt | cut -c1-160
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/aa_launches_int8_opt.csv python tools/prof_int8_step.py --opt > gpurun_out/aa_int8_opt.log 2>&1
python tools/summarize_launches.py gpurun_out/aa_launches_int8_opt.csv > gpurun_out/aa_launches_int8_opt.summary.txt 2>&1; head -22 gpurun_out/aa_launches_int8_opt.summary.txt | cut -c1-160
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_wsp -c 1 -s 2 -o gpurun_out/aa_wsp_sm1_T2048 -f python tools/prof_attn.py --T 2048 --B 32 --kind 1 > gpurun_out/aa_ncu_wsp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_one -c 1 -s 2 -o gpurun_out/aa_one_clip1_T128 -f python tools/prof_attn.py --T 128 --B 32 --kind 3 --bert-layout > gpurun_out/aa_ncu_one.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gelu_erf -c 1 -s 2 -o gpurun_out/aa_gelu -f python tools/bench_gelu.py > gpurun_out/aa_ncu_gelu.log 2>&1
for r in aa_wsp_sm1_T2048 aa_one_clip1_T128 aa_gelu; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/$r.txt 2>&1
  ncu -i gpurun_out/$r.ncu-rep --page source --csv > gpurun_out/$r.source.csv 2>/dev/null
  python tools/ncu_hot.py gpurun_out/$r.source.csv 25 >> gpurun_out/$r.txt 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -5
head -50 gpurun_out/aa_gelu.txt
