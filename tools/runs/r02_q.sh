#!/bin/bash
# round 2, call q: rays released at the end of a pass and a minimum of ready rays to start one, same box
mkdir -p gpurun_out
for v in base rel24 rel16 rel32 rel1 base; do
  for w in c4 c1 c2 c5; do
    ph=""; [ $w = c5 ] && ph="--photons 1e6"
    ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_$v.so timeout 300 python bench.py --workload $w $ph --steps 3 --warmup 2 --no-cpu-baseline \
        > gpurun_out/r02_q_${v}_$w.json 2> gpurun_out/r02_q_${v}_$w.err
    python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_q_${v}_$w.json').read()); print('$v $w', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], d['shard_check']['ok'])
except Exception as e: print('$v $w ERR', e)"
  done
done
ARTES_GPU_LIB=$PWD/build/variants/libartes_gpu_rel24.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_q_pytest.log 2>&1; tail -3 gpurun_out/r02_q_pytest.log
