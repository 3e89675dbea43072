#!/bin/bash
mkdir -p gpurun_out
timeout 600 bash tools/run_variants.sh 1024,2048,8192 1 t1 stg700 stg5000 stg30000 > gpurun_out/m_stagger.txt 2>&1
cat gpurun_out/m_stagger.txt
