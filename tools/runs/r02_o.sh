#!/bin/bash
# round 2, call o: pass statistics of the ray/event scheduler (-DE2_STATS build)
mkdir -p gpurun_out
export ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_stats.so E2_STATS=1
for w in c4 c1 c5 c2; do n=4e6; [ $w = c5 ] && n=1e6; python tools/gpu_tune.py $w $n; done > gpurun_out/r02_o_stats.txt 2>&1
cat gpurun_out/r02_o_stats.txt
