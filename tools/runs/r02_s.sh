#!/bin/bash
# round 2, call s: stress run of every scheduler (random sizes, options, detectors), bounded by timeout
mkdir -p gpurun_out
timeout 900 python tools/gpu_stress.py 400 > gpurun_out/r02_s_stress.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_s_stress.txt; tail -5 gpurun_out/r02_s_stress.txt
