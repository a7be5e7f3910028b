#!/bin/bash
# round 2, call p: which event list a warp serves -- fixed priority (pol0) against largest backlog (pol1), same box
mkdir -p gpurun_out
for v in pol0 pol1 pol0 pol1; do
  for w in c4 c1 c2 c5; do
    ph=""; [ $w = c5 ] && ph="--photons 1e6"
    ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --workload $w $ph --steps 3 --warmup 2 --no-cpu-baseline \
        > gpurun_out/r02_p_${v}_$w.json 2> gpurun_out/r02_p_${v}_$w.err
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_p_${v}_$w.json').read()); print('$v $w', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'])
except Exception as e: print('$v $w ERR', e)"
  done
done
export ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_pol1stats.so E2_STATS=1
for w in c4 c1 c5 c2; do n=4e6; [ $w = c5 ] && n=1e6; python tools/gpu_tune.py $w $n; done > gpurun_out/r02_p_stats.txt 2>&1
cat gpurun_out/r02_p_stats.txt
